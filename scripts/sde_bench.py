#!/usr/bin/env python3
"""Config 5 timing (development tool): van der Pol SDE population N = 2^22, Euler-Maruyama with keyed Philox noise, tau-leaping
birth / death sweeps (scan compaction).  Euler roofline: 48 B per particle-step (read x,y 16 B + write x,y 16 B + rates
16 B; the rate columns are only written when outputs are due, so a plain step moves 32 B + the 4-byte id column the noise is keyed by)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "vdp_sde.pop"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
pop = capi.Population(lib, capacity=2 * n)
pop.init(n, replay=False, seed=8)
t = pop.run_fixed("fe", 0.0, 0.001, 20); pop.sync()
best = 1e9
for rep in range(3):
    pop.event_record(0); t = pop.run_fixed("fe", t, 0.001, 200); pop.event_record(1)
    best = min(best, pop.event_elapsed_ms(0, 1) / 200)
print("vdp_sde fe  N=%d  %.4f ms/step  %.3e particle-steps/s  %.0f GB/s at 48 B/particle-step = %.3f of 6550 GB/s" % (
    n, best, n / (best * 1e-3), 48.0 * n / (best * 1e-3) / 1e9, 48.0 * n / (best * 1e-3) / 6550.4e9), flush=True)
# tau-leap sweeps (each: rates -> flags -> scans -> spawn generations -> compaction -> event rows)
pop.rhs(t); pop.sync()
for k in range(6):
    n0 = pop.size()[0]
    pop.sync()
    w0 = time.perf_counter()
    pop.event_record(0)
    rows = pop.tauleap_apply(t, 1.0, k, as_array=True, max_rows=(1 << 21) if os.environ.get('SDE_COPY_ROWS') else None)
    pop.event_record(1)
    pop.sync()
    w = time.perf_counter() - w0
    ms = pop.event_elapsed_ms(0, 1)
    n1 = pop.size()[0]
    print("tau-leap sweep %d: N %d -> %d, %d event rows, %.3f ms on the stream (%.2f ms wall incl. D2H of the rows) = %.3e particles/s" % (
        k, n0, n1, len(rows), ms, 1e3 * w, n0 / (ms * 1e-3)), flush=True)
    pop.rhs(t)
pop.close()
