"""GPU parity for the low-rank %C coupling (config 3) against the CPU restatement of src/lrexp.c (oracle/lowrank.c,
itself bit-identical to the unchanged reference binary on this fixture, tests/test_oracle_port.py).

Tolerance: the expansion converts a Chebyshev fit to the monomial basis (integer coefficients up to ~1e11 at the order
selected here, alternating signs), so round-off differences between the serial CPU sums and the blocked GPU sums are
amplified; psi agrees to 1e-8 relative to its scale, the state to 1e-10."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("meth,ps,ts,res", [("dde", 256, (0, 0.4), 3), ("fe", 1000, (0, 0.2), 2), ("rkck", 300, (0, 1.0), 3)])
def test_lowrank_matches_restatement(model_lib, meth, ps, ts, res):
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib("kuramoto_lowrank")
    opts = {"meth": meth, "h0": 0.05 if meth == "fe" else 0.1, "res": res, "ps": ps}
    pop = capi.Population(lib, capacity=max(1024, ps))
    out = capi.odesolver(pop, ts, meth=meth, res=res, popsize=ps, h0=opts["h0"], replay_ic=True)
    pop.close()
    ref = port.run("kuramoto_lowrank", opts, ts)
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb)
    np.testing.assert_array_equal(ta["id"], tb["id"])
    first_a, first_b = ta[ta["step"] == 0]["cols"], tb[tb["step"] == 0]["cols"]
    np.testing.assert_array_equal(first_a[:, 0], first_b[:, 0])              # theta0: replayed libc stream
    np.testing.assert_array_equal(first_a[:, 6:], first_b[:, 6:])            # omega, U[0:8], V[0:8]
    psi_a, psi_b = ta["cols"][:, 3], tb["cols"][:, 3]
    scale = np.abs(psi_b).max()
    print("max |dpsi|/scale", np.abs(psi_a - psi_b).max() / scale, "max rel dtheta", (np.abs(ta["cols"][:, 0] - tb["cols"][:, 0]) / np.abs(tb["cols"][:, 0])).max())
    assert np.abs(psi_a - psi_b).max() <= 1e-8 * scale
    np.testing.assert_allclose(ta["cols"][:, 0], tb["cols"][:, 0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-11)


def test_lowrank_matches_direct_sum(model_lib):
    """independent check of the expansion itself: psi_i vs the O(N^2) sum_j (U_i.V_j) sin(x_j - x_i) in numpy"""
    from odexp_b200 import capi
    lib = model_lib("kuramoto_lowrank")
    n = 512
    pop = capi.Population(lib, capacity=1024)
    pop.init(n, replay=True)
    pop.rhs(0.0)
    d = pop.download()
    x, U, V, psi = d["y"][:, 0], d["expr"][:, 1:9], d["expr"][:, 9:17], d["psi"][:, 0]
    direct = ((U @ V.T) * np.sin(x[None, :] - x[:, None])).sum(axis=1)
    pop.close()
    # lrabstol / lrreltol = 1e-6 of the reference bound the truncation error of the expansion
    assert np.abs(psi - direct).max() <= 1e-4 * np.abs(direct).max()


@pytest.mark.parametrize("ps", [203, 3001, 8192])
def test_lowrank_rank64_matches_restatement(model_lib, ps):
    """the rank-64 kernels of BASELINE config 3 (TMA-staged moment / apply contractions) at ragged sizes: fewer particles than
    one tile, an odd count, whole tiles; the restatement is pinned to the reference binary for this fixture (test_oracle_port)"""
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib("kuramoto_lowrank64")
    opts = {"meth": "dde", "h0": 0.1, "res": 3, "ps": ps}
    pop = capi.Population(lib, capacity=max(1024, ps))
    out = capi.odesolver(pop, (0, 0.4), meth="dde", res=3, popsize=ps, h0=0.1, replay_ic=True)
    pop.close()
    ref = port.run("kuramoto_lowrank64", opts, (0, 0.4))
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb)
    np.testing.assert_array_equal(ta["id"], tb["id"])
    np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, 6:], tb[tb["step"] == 0]["cols"][:, 6:])     # omega, U, V
    psi_a, psi_b = ta["cols"][:, 3], tb["cols"][:, 3]
    scale = np.abs(psi_b).max()
    print("max |dpsi|/scale", np.abs(psi_a - psi_b).max() / scale)
    assert np.abs(psi_a - psi_b).max() <= 1e-8 * scale
    np.testing.assert_allclose(ta["cols"][:, 0], tb["cols"][:, 0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("meth,ps,ts,res", [("dde", 100, (0, 2.0), 5), ("rkf45", 57, (0, 25.0), 4), ("fe", 1500, (0, 12.0), 3)])
def test_lowrank_in_an_auxiliary_variable(model_lib, meth, ps, ts, res):
    """`%C CPL lrwexp {var = theta}` with theta an AUXILIARY variable (the wrapped phase), as in the reference's shipped
    examples/kuramoto_lowrank.pop: the gather reads the aux values pass 1 of pop_rhs left (mvar, src/datastruct.c:824-855); plus
    MA() inside %M.  The long runs take the phases past 2 pi, where the wrapped variable jumps."""
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib("kuramoto_lowrank_aux")
    opts = {"meth": meth, "h0": 0.1, "res": res, "ps": ps}
    pop = capi.Population(lib, capacity=max(1024, ps))
    out = capi.odesolver(pop, ts, meth=meth, res=res, popsize=ps, h0=0.1, replay_ic=True)
    pop.close()
    ref = port.run("kuramoto_lowrank_aux", opts, ts)
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb)
    np.testing.assert_array_equal(ta["id"], tb["id"])
    np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, 6:], tb[tb["step"] == 0]["cols"][:, 6:])     # omega, U, V
    if ts[1] > 10:
        assert (tb["cols"][:, 0] > 2 * 3.141592).any() and (tb["cols"][:, 2] < tb["cols"][:, 0] - 6).any(), "phases must wrap"
    psi_a, psi_b = ta["cols"][:, 3], tb["cols"][:, 3]
    scale = np.abs(psi_b).max()
    print("max |dpsi|/scale", np.abs(psi_a - psi_b).max() / scale)
    assert np.abs(psi_a - psi_b).max() <= 1e-8 * scale
    np.testing.assert_allclose(ta["cols"][:, 0], tb["cols"][:, 0], rtol=1e-10, atol=1e-12)            # theta
    np.testing.assert_allclose(ta["cols"][:, 1:3], tb["cols"][:, 1:3], rtol=0, atol=1e-9)             # aux y, aux theta (wrapped)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("meth,ps,ts", [("dde", 200, (0, 1.0)), ("rkck", 77, (0, 2.0)), ("fe", 5000, (0, 0.3))])
def test_unweighted_lrexp(model_lib, meth, ps, ts):
    """`%C CPL lrexp {var, fun}` = lrkern (src/lrexp.c:484-496): the rank-1 kernels with a column of ones, against the
    restatement of lrexp() (bit-identical to the reference binary) and against the direct O(N^2) sum"""
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib("kuramoto_lrexp")
    opts = {"meth": meth, "h0": 0.1, "res": 3, "ps": ps}
    pop = capi.Population(lib, capacity=max(1024, ps))
    out = capi.odesolver(pop, ts, meth=meth, res=3, popsize=ps, h0=0.1, replay_ic=True)
    pop.close()
    ref = port.run("kuramoto_lrexp", opts, ts)
    ta, tb = out["traj"], ref["traj"]                       # cols: theta, aux cpl, CPL, A1, A2, omega
    assert len(ta) == len(tb)
    psi_a, psi_b = ta["cols"][:, 2], tb["cols"][:, 2]
    scale = np.abs(psi_b).max()
    print("max |dpsi|/scale", np.abs(psi_a - psi_b).max() / scale)
    assert np.abs(psi_a - psi_b).max() <= 1e-8 * scale
    np.testing.assert_allclose(ta["cols"][:, 0], tb["cols"][:, 0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-11)
    # the expansion itself against the direct sum, on the state the RHS was evaluated at
    pop = capi.Population(lib, capacity=max(1024, ps))
    pop.init(ps, replay=True)
    pop.rhs(0.0)
    d = pop.download()
    pop.close()
    th = d["y"][:, 0]
    direct = np.sin(th[None, :] - th[:, None]).sum(axis=1)
    assert np.abs(d["psi"][:, 0] - direct).max() <= 1e-4 * max(1.0, np.abs(direct).max())      # lrabstol / lrreltol = 1e-6
