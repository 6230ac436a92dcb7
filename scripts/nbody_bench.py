#!/usr/bin/env python3
"""N-body %C pairwise kernel timing (development tool): time per RHS evaluation, pair rate, FP64 roofline fraction."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "nbody.pop"))
for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "32768,131072").split(",")]:
    pop = capi.Population(lib, capacity=n)
    pop.init(n, replay=False)
    pop.rhs(0.0); pop.sync()
    reps = 3
    pop.event_record(0)
    for _ in range(reps):
        pop.rhs(0.0)
    pop.event_record(1)
    ms = pop.event_elapsed_ms(0, 1) / reps
    pairs = float(n) * n
    print("N=%d  %.3f ms per RHS evaluation  %.3e pairs/s  %.2f TFLOP/s at 20 flop/pair (%.1f%% of 37.2 nominal)" % (
        n, ms, pairs / (ms * 1e-3), 20 * pairs / (ms * 1e-3) / 1e12, 100 * 20 * pairs / (ms * 1e-3) / 37.2e12))
    if n <= 131072:
        pop.set_tolerances(1e-8, 1e-8)
        t, h = 0.0, 1e-3
        pop.evolve_reset()
        w0 = time.perf_counter(); steps = 0
        while steps < 3:
            t, h = pop.evolve_apply("rkf45", t, 1.0, h); steps += 1
        pop.sync()
        w = time.perf_counter() - w0
        print("   rkf45: %d accepted steps, %.1f ms/step (attempts incl.), %.3e particle-steps/s, stats %s" % (steps, 1e3 * w / steps, n * steps / w, pop.evolve_stats()))
    pop.close()
