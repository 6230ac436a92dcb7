#!/usr/bin/env python3
"""Development check of the lean RK4 scheme (option "lean"): states after a few steps against the stored scheme at several
sizes (ragged ones included), then time per step at the BASELINE size."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "kuramoto.pop"))


def state(n, steps, tma, lean):
    pop = capi.Population(lib, capacity=n)
    pop.set_option("tma", tma); pop.set_option("lean", lean)
    pop.init(n, replay=False)
    pop.run_fixed("dde", 0.0, 0.05, steps)
    pop.sync()
    y = pop.download()["y"][:, 0].copy(); m = np.array(pop.meanfields())
    pop.close()
    return y, m


for n, steps in ((1000, 40), (4097, 7), (235931, 5), (1 << 21, 3)):
    y0, m0 = state(n, steps, 0, 0)
    y1, m1 = state(n, steps, 1, 0)
    y2, m2 = state(n, steps, 1, 1)
    print("N=%d steps=%d  ring-stored vs plain: max|d|=%.3e   lean vs plain: max rel=%.3e  mfd diff=%.3e" % (
        n, steps, np.abs(y1 - y0).max(), (np.abs(y2 - y0) / np.maximum(1.0, np.abs(y0))).max(), np.abs(m2 - m0).max()), flush=True)

n = 1 << 24
for lean in (0, 1):
    pop = capi.Population(lib, capacity=n)
    pop.set_option("lean", lean)
    pop.init(n, replay=False)
    t = pop.run_fixed("dde", 0.0, 0.05, 10); pop.sync()
    best = 1e9
    for rep in range(3):
        pop.event_record(0); t = pop.run_fixed("dde", t, 0.05, 100); pop.event_record(1)
        best = min(best, pop.event_elapsed_ms(0, 1) / 100)
    print("N=2^24 lean=%d  %.4f ms/step  %.3e particle-steps/s  frac72=%.3f  mfd=%s" % (lean, best, n / (best * 1e-3), 72.0 * n / (best * 1e-3) / 6550.4e9, pop.meanfields()), flush=True)
    pop.close()
