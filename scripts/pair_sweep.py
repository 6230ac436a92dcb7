#!/usr/bin/env python3
"""N-body %C pair kernel: compile-time variants (development tool).  `build` compiles them here; without arguments it times them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import build, capi
VARIANTS = {
    "nbody": [],
    "nbody_u4": ["-DPAIR_UNROLL_N=4"],
    "nbody_u1": ["-DPAIR_UNROLL_N=1"],
    "nbody_t256": ["-DPAIR_THREADS_N=256"],
    "nbody_t64": ["-DPAIR_THREADS_N=64"],
    "nbody_tj256": ["-DPAIR_TJ_N=256"],
    "nbody_w4": ["-DPAIR_WAVES=4"],
    "nbody_w2": ["-DPAIR_WAVES=2"],
}
text = open(os.path.join(ROOT, "models", "nbody.pop"), encoding="utf8").read()
if len(sys.argv) > 1 and sys.argv[1] == "build":
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as ex:
        list(ex.map(lambda kv: build.build_model(text=text, name=kv[0], extra_flags=kv[1]), VARIANTS.items()))
    sys.exit(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
for name in VARIANTS:
    lib = capi.ModelLib(build.lib_path(name))
    pop = capi.Population(lib, capacity=n)
    pop.init(n, replay=False)
    pop.rhs(0.0); pop.sync()
    best = 1e30
    for _ in range(3):
        pop.event_record(0)
        for _ in range(2):
            pop.rhs(0.0)
        pop.event_record(1)
        best = min(best, pop.event_elapsed_ms(0, 1) / 2)
    pairs = float(n) * n
    print("%-12s N=%d  %.3f ms per evaluation  %.3e pairs/s  %.1f%% of 37.2 TFLOP/s at 20 flop/pair" % (name, n, best, pairs / (best * 1e-3), 100 * 20 * pairs / (best * 1e-3) / 37.2e12), flush=True)
    pop.close()
