#!/usr/bin/env python3
"""run a short persistent Kuramoto RK4 burst of a kernel-variant library (scripts/kbench.py build) at a given N (for ncu):
   python scripts/prof_kur.py <tag> [N] [steps] [option=value ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi, build as b
tag = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lib = capi.ModelLib(b.lib_path("kuramoto_" + tag))
pop = capi.Population(lib, capacity=n)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    pop.set_option(k, float(v))
pop.init(n, replay=False)
t = pop.run_fixed("dde", 0.0, 0.05, 5)
pop.sync()
pop.event_record(0); t = pop.run_fixed("dde", t, 0.05, steps); pop.event_record(1)
print("ms/step", pop.event_elapsed_ms(0, 1) / steps)
pop.close()
