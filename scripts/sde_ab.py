#!/usr/bin/env python3
"""Euler-Maruyama step time of variant libraries of the SDE fixture (development tool): python scripts/sde_ab.py tag[:option=value,...] ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi, build
n = 1 << 22
for arg in sys.argv[1:]:
    tag, _, opts = arg.partition(":")
    lib = capi.ModelLib(build.lib_path("vdp_sde_" + tag))
    pop = capi.Population(lib, capacity=2 * n)
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("=")
        pop.set_option(k, float(v))
    pop.init(n, replay=False, seed=8)
    t = pop.run_fixed("fe", 0.0, 0.001, 20); pop.sync()
    best = 1e9
    for rep in range(3):
        pop.event_record(0); t = pop.run_fixed("fe", t, 0.001, 200); pop.event_record(1)
        best = min(best, pop.event_elapsed_ms(0, 1) / 200)
    y = pop.download(0, 1000)["y"]
    print("%s: %.4f ms/step  %.3e particle-steps/s  %.3f of 6550 GB/s at 48 B  y[0]=%r" % (arg, best, n / (best * 1e-3), 48.0 * n / (best * 1e-3) / 6550.4e9, y[0].tolist()), flush=True)
    pop.close()
