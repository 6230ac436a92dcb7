#!/usr/bin/env python3
"""tau-leaping sweep timing (development tool): device time between events and host wall time per sweep at N = 2^22"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
from odexp_b200 import build as _b
lib = capi.ModelLib(_b.lib_path("vdp_sde_timing")) if os.environ.get("LEAP_TIMING") else capi.ModelLib.from_pop(os.path.join(ROOT, "models", "vdp_sde.pop"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pop = capi.Population(lib, capacity=2 * n + 8192)
pop.init(n, replay=False, seed=8)
t = pop.run_fixed("fe", 0.0, 0.001, 20); pop.sync()
pop.rhs(t); pop.sync()
for k in range(nsw):
    n0 = pop.size()[0]
    pop.sync()
    w0 = time.perf_counter()
    pop.event_record(0)
    rows = pop.tauleap_apply(t, 1.0, k, as_array=True)
    pop.event_record(1)
    pop.sync()
    w = time.perf_counter() - w0
    print("sweep %d: N %d -> %d, %d rows, %.3f ms between events, %.3f ms wall" % (k, n0, pop.size()[0], len(rows), pop.event_elapsed_ms(0, 1), 1e3 * w), flush=True)
    pop.rhs(t)
pop.close()
