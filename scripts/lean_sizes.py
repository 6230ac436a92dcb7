#!/usr/bin/env python3
"""Development tool: ms per RK4 step (scaled to N = 2^24) of the persistent kernel variants at L2-resident and HBM-resident sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "kuramoto.pop"))
for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1048576,2097152,4194304,8388608,16777216,33554432").split(",")]:
    for tma, lean in ((0, 0), (1, 0), (1, 1)):
        pop = capi.Population(lib, capacity=n)
        pop.set_option("tma", tma); pop.set_option("lean", lean)
        pop.init(n, replay=False)
        steps = max(20, int(100 * (1 << 24) / n))
        t = pop.run_fixed("dde", 0.0, 0.05, 10); pop.sync()
        best = 1e9
        for rep in range(3):
            pop.event_record(0); t = pop.run_fixed("dde", t, 0.05, steps); pop.event_record(1)
            best = min(best, pop.event_elapsed_ms(0, 1) / steps)
        print("N=%9d tma=%d lean=%d  %.4f ms/step  = %.4f ms per 2^24 particles  %.3e particle-steps/s" % (n, tma, lean, best, best * (1 << 24) / n, n / (best * 1e-3)), flush=True)
        pop.close()
