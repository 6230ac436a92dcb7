#!/usr/bin/env python3
"""Symmetric vs ordered all-pairs %C kernel (development tool): ms per RHS evaluation at N = 2^17 for register-blocking variants.
Build the variant libraries where nvcc is (`--build`), run on the GPU box without arguments."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import capi, build
VARIANTS = {"p82b1": (8, 2, 1), "p82b3": (8, 2, 3), "p81b3": (8, 1, 3), "p44b2": (4, 4, 2)}
if os.environ.get("PSYM_ALL"):
    VARIANTS.update({"p44b1": (4, 4, 1), "p42b2": (4, 2, 2), "p84b1": (8, 4, 1), "p22b4": (2, 2, 4), "p81b2": (8, 1, 2)})
pop_path = os.path.join(ROOT, "models", "nbody.pop")
if "--build" in sys.argv:
    for name, (pi, pj, mb) in VARIANTS.items():
        print(build.build_model(pop_path, name="nbody_" + name, extra_flags=("-DPSYM_PI=%d" % pi, "-DPSYM_PJ=%d" % pj, "-DPSYM_MINB=%d" % mb)))
    sys.exit(0)
n = 1 << 17
for name in VARIANTS:
    so = build.lib_path("nbody_" + name)
    if not os.path.exists(so):
        continue
    lib = capi.ModelLib(so)
    pop = capi.Population(lib, capacity=n)
    pop.init(n, replay=False)
    for sym in ((0, 1) if name == "p82b1" else (1,)):
        pop.set_option("pair_sym", sym)
        pop.rhs(0.0); pop.sync()
        reps = 5
        pop.event_record(0)
        for _ in range(reps):
            pop.rhs(0.0)
        pop.event_record(1)
        ms = pop.event_elapsed_ms(0, 1) / reps
        pairs = float(n) * n
        print("%s sym=%d N=%d  %.3f ms per RHS evaluation  %.3e pairs/s  %.2f TFLOP/s at 20 flop/pair (%.3f of 36.94 measured)" % (
            name, sym, n, ms, pairs / (ms * 1e-3), 20 * pairs / (ms * 1e-3) / 1e12, 20 * pairs / (ms * 1e-3) / 36.94e12), flush=True)
    pop.close()
