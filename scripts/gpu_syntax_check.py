"""GPU side of the model-language additions of round 2 (arrays with several indices; names that clash with C++, max() / min() / abs()):
the two models of tests/test_emitter_vs_reference.py::SYNTAX_MODELS through the C-ABI on the GPU (odx_odesolver, replayed initial
conditions, dde and fe) against outputs of the unchanged reference binary (tests/golden/syntax_<model>_<meth>.npz, written by this script
with --make where /root/reference exists).  No torch: ctypes + numpy only, a few seconds on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_emitter_vs_reference import SYNTAX_MODELS      # noqa: E402
from odexp_b200 import build, capi                         # noqa: E402

from test_gpu_syntax import CASES                         # noqa: E402
OPTS = dict(h0=0.05, res=4, ps=7)
GOLD = os.path.join(ROOT, "tests", "golden")

if "--make" in sys.argv:
    from oracle import build_ref, refrun
    assert build_ref.build()
    for k, meth in CASES:
        text, dims, _ = SYNTAX_MODELS[k]
        build_ref.build_one("syntax_" + k, text)
        ref = refrun.run("syntax_" + k, dict(OPTS, meth=meth, hmin=1e-5, abstol=1e-6, reltol=1e-6, ssahmin=0.0, aleap=0.01, seed=3141592), (0, 0.5), dims=dims)
        np.savez_compressed(os.path.join(GOLD, "syntax_%s_%s.npz" % (k, meth)), traj=ref["traj"], mfd=ref["stats"]["mfd"])
        build.build_model(text=text, name="syntax_" + k)          # the device library travels to the GPU box prebuilt
    import shutil
    for k in sorted({c[0] for c in CASES}):
        shutil.rmtree(os.path.join(build_ref.OUT, "syntax_" + k), ignore_errors=True)
    print("goldens and libraries written")
    sys.exit(0)

from test_gpu_syntax import check_case                    # noqa: E402

worst = 0.0
for k, meth in CASES:
    err, merr, rows = check_case(k, meth)
    worst = max(worst, err, merr)
    print("%-15s %-4s rows %3d  max rel err %.2e  mean fields %.2e" % (k, meth, rows, err, merr), flush=True)
    assert err <= 1e-10 and merr <= 1e-12, (k, meth, err, merr)
print("GPU_SYNTAX_CHECK OK worst %.2e" % worst)
