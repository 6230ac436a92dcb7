#!/usr/bin/env python3
"""step-kernel variant crossover (development tool): ms per RK4 step of the plain / ring-stored / ring-lean persistent kernels
at several population sizes on one GPU:  python scripts/variant_matrix.py [tag]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi, build as b
tag = sys.argv[1] if len(sys.argv) > 1 else None
lib = capi.ModelLib(b.lib_path("kuramoto_" + tag)) if tag else capi.ModelLib.from_pop(os.path.join(ROOT, "models", "kuramoto.pop"))
for n in (1 << 21, 3 << 20, 1 << 22, 6 << 20, 1 << 23, 1 << 24):
    row = []
    for name, opts in (("plain", {"tma": 0}), ("ring-stored", {"tma": 1, "lean": 0}), ("ring-lean", {"tma": 1, "lean": 1}), ("auto", {})):
        pop = capi.Population(lib, capacity=n)
        for k, v in opts.items():
            pop.set_option(k, v)
        pop.init(n, replay=False)
        steps = 100 if n >= (1 << 23) else 300
        t = pop.run_fixed("dde", 0.0, 0.05, 10); pop.sync()
        best = 1e9
        for rep in range(3):
            pop.event_record(0); t = pop.run_fixed("dde", t, 0.05, steps); pop.event_record(1)
            best = min(best, pop.event_elapsed_ms(0, 1) / steps)
        row.append("%s %.4f" % (name, best))
        pop.close()
    print("N=%9d  " % n + "   ".join(row), flush=True)
