#!/usr/bin/env python3
"""config 3 timing: kuramoto_lowrank64 (rank 64), N = 2^20, one RHS evaluation and RK4 steps"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "kuramoto_lowrank64.pop"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
pop = capi.Population(lib, capacity=n)
pop.init(n, replay=False)
pop.rhs(0.0); pop.sync()
pop.event_record(0)
for _ in range(5):
    pop.rhs(0.0)
pop.event_record(1)
ms = pop.event_elapsed_ms(0, 1) / 5
gb = (2 * n * 64 * 8 + 2 * n * 8) / 1e9
print("N=%d rank 64: %.3f ms per RHS evaluation (U,V stream %.2f GB -> %.0f GB/s of algorithmic bytes)" % (n, ms, gb, gb / (ms * 1e-3)))
pop.event_record(0); t = pop.run_fixed("dde", 0.0, 0.05, 5); pop.event_record(1)
ms = pop.event_elapsed_ms(0, 1) / 5
print("RK4: %.3f ms/step  %.3e particle-steps/s" % (ms, n / (ms * 1e-3)))
d = pop.download(0, 4)
print("psi sample", d["psi"][:, 0], "mfd", pop.meanfields())
pop.close()
