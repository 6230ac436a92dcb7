#!/usr/bin/env python3
"""short Euler-Maruyama burst of the SDE fixture at N = 2^22 for ncu: python scripts/prof_sde.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
n = 1 << 22
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "vdp_sde.pop"))
pop = capi.Population(lib, capacity=2 * n)
pop.init(n, replay=False, seed=8)
t = pop.run_fixed("fe", 0.0, 0.001, 20); pop.sync()
pop.event_record(0); t = pop.run_fixed("fe", t, 0.001, steps); pop.event_record(1)
print("ms/step", pop.event_elapsed_ms(0, 1) / steps)
pop.close()
