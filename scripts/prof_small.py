#!/usr/bin/env python3
"""run a short persistent Kuramoto RK4 burst at a given N (for ncu)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "kuramoto.pop"))
pop = capi.Population(lib, capacity=n)
pop.init(n, replay=False)
t = pop.run_fixed("dde", 0.0, 0.05, 5)
pop.sync()
pop.event_record(0); t = pop.run_fixed("dde", t, 0.05, steps); pop.event_record(1)
print("ms/step", pop.event_elapsed_ms(0, 1) / steps)
pop.close()
