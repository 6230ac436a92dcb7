#!/usr/bin/env python3
"""two RHS evaluations of the N-body fixture at N (default 2^17) for ncu: python scripts/prof_nbody.py [N] [pair_sym]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "nbody.pop"))
pop = capi.Population(lib, capacity=n)
if len(sys.argv) > 2:
    pop.set_option("pair_sym", float(sys.argv[2]))
pop.init(n, replay=False)
pop.rhs(0.0); pop.sync()
pop.event_record(0); pop.rhs(0.0); pop.event_record(1)
print("ms per evaluation", pop.event_elapsed_ms(0, 1))
pop.close()
