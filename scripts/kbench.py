#!/usr/bin/env python3
"""Kernel-variant micro benchmark for the Kuramoto RK4 step (development tool).
  build (CPU box):  python scripts/kbench.py build
  run (GPU box):    python scripts/kbench.py run
Variants are compile-time flags of odx_engine.cuh; each becomes odexp_b200/_build/libodx_kuramoto_<tag>.so."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = json.loads(os.environ.get("KBENCH_VARIANTS", '{"mb2": ["-DCHAIN_MIN_BLOCKS=2"], "mb3": ["-DCHAIN_MIN_BLOCKS=3"], "mb4": ["-DCHAIN_MIN_BLOCKS=4"]}'))
POP = os.path.join(ROOT, "models", "kuramoto.pop")


def build():
    from odexp_b200 import build as b
    for tag, flags in VARIANTS.items():
        print(tag, b.build_model(POP, name="kuramoto_" + tag, extra_flags=flags))


def run(n=1 << 24, steps=100):
    from odexp_b200 import capi, build as b
    for tag in VARIANTS:
        lib = capi.ModelLib(b.lib_path("kuramoto_" + tag))
        for persistent in ((1, 0) if os.environ.get("KBENCH_BOTH") else (1,)):
            pop = capi.Population(lib, capacity=n)
            pop.set_option("persistent", persistent)
            pop.init(n, replay=False)
            t = pop.run_fixed("dde", 0.0, 0.05, 10)
            pop.sync()
            best = 1e9
            for rep in range(3):
                pop.event_record(0); t = pop.run_fixed("dde", t, 0.05, steps); pop.event_record(1)
                best = min(best, pop.event_elapsed_ms(0, 1) / steps)
            print("%-8s persistent=%d %.4f ms/step  %.3e particle-steps/s  frac72=%.3f  mfd=%s" % (
                tag, persistent, best, n / (best * 1e-3), 72.0 * n / (best * 1e-3) / 6550.4e9, pop.meanfields()))
            pop.close()


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        for n in [int(x) for x in os.environ.get("KBENCH_N", str(1 << 24)).split(",")]:
            print("N =", n)
            run(n=n, steps=100 if n >= (1 << 22) else 400)
