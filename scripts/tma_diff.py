#!/usr/bin/env python3
"""Development check: persistent-kernel variants must give bit-identical states (same arithmetic, same order)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi, build as b

def run(tag, n, steps, persistent=1):
    lib = capi.ModelLib(b.lib_path("kuramoto_" + tag))
    pop = capi.Population(lib, capacity=n)
    pop.set_option("persistent", persistent)
    pop.init(n, replay=False)
    pop.run_fixed("dde", 0.0, 0.05, steps)
    pop.sync()
    y = pop.download()["y"][:, 0].copy()
    pop.close()
    return y

n = int(os.environ.get("N", 1 << 24)); steps = int(os.environ.get("STEPS", 1))
ref = run(sys.argv[1], n, steps)
for tag in sys.argv[2:]:
    for rep in range(2):
        y = run(tag, n, steps)
        bad = np.nonzero(y != ref)[0]
        print(tag, "rep", rep, "mismatches:", len(bad), "max|d|", np.abs(y - ref).max() if len(bad) else 0.0)
        if len(bad):
            print("  first", bad[:12], " last", bad[-5:])
            d = np.diff(bad); runs = np.split(bad, np.nonzero(d > 1)[0] + 1)
            print("  runs:", len(runs), [(int(r[0]), len(r)) for r in runs[:12]])
