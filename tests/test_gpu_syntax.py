"""The model-language additions of round 2 on the GPU: arrays with several indices, names that clash with C++, the reference's max() /
min() macros and C's abs() -- the models of tests/test_emitter_vs_reference.py::SYNTAX_MODELS through the C-ABI (odx_odesolver, replayed
initial conditions) against outputs of the unchanged reference binary (tests/golden/syntax_<model>_<meth>.npz, written by
`scripts/gpu_syntax_check.py --make`).  The CPU side of the same models: test_documented_syntax_forms_match_the_reference_binary."""
import os

import numpy as np
import pytest

from conftest import ROOT

CASES = [(k, m) for k in ("multi_index", "reserved_names", "oid_coupling") for m in ("dde", "fe")]
OPTS = dict(h0=0.05, res=4, ps=7)


def check_case(which, meth):
    from test_emitter_vs_reference import SYNTAX_MODELS
    from odexp_b200 import build, capi
    text, dims, _ = SYNTAX_MODELS[which]
    lib = capi.ModelLib(build.build_model(text=text, name="syntax_" + which))
    pop = capi.Population(lib, capacity=16)
    out = capi.odesolver(pop, (0, 0.5), meth=meth, res=OPTS["res"], popsize=OPTS["ps"], h0=OPTS["h0"], replay_ic=True)
    pop.close()
    g = np.load(os.path.join(ROOT, "tests", "golden", "syntax_%s_%s.npz" % (which, meth)))
    ta, tb = out["traj"], g["traj"]
    assert len(ta) == len(tb) and np.array_equal(ta["id"], tb["id"])
    first = ta["step"] == 0
    nv, ne = dims["nvar"], dims["nexpr"]
    assert np.array_equal(ta["cols"][first][:, :nv], tb["cols"][first][:, :nv])          # replayed initial conditions: bit-identical
    assert np.array_equal(ta["cols"][:, -ne:], tb["cols"][:, -ne:])                      # parametric expressions
    scale = np.maximum(np.abs(tb["cols"]).max(axis=0), 1e-6)
    return (np.abs(ta["cols"] - tb["cols"]) / scale).max(), np.abs(out["stats"]["mfd"] - g["mfd"]).max(), len(ta)


@pytest.mark.gpu
@pytest.mark.parametrize("which,meth", CASES)
def test_syntax_models_on_the_gpu_match_the_reference_binary(which, meth):
    err, merr, rows = check_case(which, meth)
    assert rows == OPTS["ps"] * OPTS["res"]
    assert err <= 1e-10 and merr <= 1e-12, (err, merr)         # measured on a B200: 4e-16 / 1e-16 (profiles/r2g_gpu_syntax_check.txt)
