"""Symmetric all-pairs %C kernel (k_pair_sym, Newton's third law) against the ordered kernel k_pair, which is the one the
oracle / reference-binary parity tests pin (tests/test_gpu_parity_golden.py).  The symmetric kernel evaluates every unordered
pair once; its sums run in another (fixed) order, so psi agrees with the ordered kernel up to rounding, not bit for bit.
Reference semantics: src/odexp:905-954 (psi_k(i) = (1/N) sum_j expr_k(i, j), self term guarded by the model)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _psi(pop, sym, t=0.0):
    pop.set_option("pair_sym", sym)
    pop.rhs(t)
    return pop.download()["psi"].copy()


@pytest.mark.parametrize("n", [2, 31, 128, 129, 1023, 1024, 1025, 4099, 5 * 1024, 20000])
def test_symmetric_pairs_match_ordered_kernel(model_lib, n):
    """full, ragged and tiny populations (one tile, several tiles, a last tile of one particle): forced symmetric kernel vs
    ordered kernel, relative to the largest |psi| of the component (1e-12: the two sums differ by the order of ~n additions)"""
    from odexp_b200 import capi
    pop = capi.Population(model_lib("nbody"), capacity=n)
    pop.init(n, replay=False, seed=11)
    ordered = _psi(pop, 0)
    sym = _psi(pop, 1)
    assert np.all(np.isfinite(sym))
    scale = np.abs(ordered).max(axis=0)
    assert np.all(np.abs(sym - ordered).max(axis=0) <= 1e-12 * scale), (np.abs(sym - ordered).max(axis=0) / scale)
    again = _psi(pop, 1)
    np.testing.assert_array_equal(sym, again)          # fixed schedule, no atomics: bit-reproducible
    pop.close()


def test_symmetric_pairs_against_numpy(model_lib):
    """direct O(N^2) sum in numpy (the model's expression, examples/Nbody.pop:12-18) at a size with several tiles"""
    from odexp_b200 import capi
    n = 3000
    pop = capi.Population(model_lib("nbody"), capacity=n)
    pop.init(n, replay=False, seed=5)
    pos = pop.download()["y"][:, :3]
    sym = _psi(pop, 1)
    d = pos[:, None, :] - pos[None, :, :]
    r2 = (d * d).sum(axis=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(r2 > 0, r2 ** -1.5, 0.0)
    ref = -6.674e-11 * 1e9 * (d * w[:, :, None]).sum(axis=1) / n
    np.testing.assert_allclose(sym, ref, rtol=0, atol=1e-11 * np.abs(ref).max())
    pop.close()


def test_symmetric_pairs_coincident_particles_take_the_exact_pass(model_lib):
    """two distinct particles at the same position: the fast term is 0 x inf = NaN; the finish kernel raises the flag and the
    ordered kernels redo the evaluation with the guarded form -- the result is the ordered kernel's, bit for bit"""
    from odexp_b200 import capi
    n = 2500
    pop = capi.Population(model_lib("nbody"), capacity=n)
    pop.init(n, replay=False, seed=7)
    st = pop.download()
    y = st["y"].copy()
    y[1700, :3] = y[40, :3]
    pop.upload(y, st["expr"])
    ordered = _psi(pop, 0)
    sym = _psi(pop, 1)
    assert np.all(np.isfinite(ordered))
    np.testing.assert_array_equal(sym, ordered)
    pop.close()


def test_symmetric_pairs_rkf45_trajectory(model_lib):
    """whole adaptive run (rkf45, 6 evaluations per step) with the symmetric kernel vs the ordered one: same accepted / rejected
    step sequence and positions to 1e-10 over a short horizon (the N-body problem is chaotic: SURVEY 8(d) config 4)"""
    from odexp_b200 import capi
    n = 4096
    res = []
    for sym in (0, 1):
        pop = capi.Population(model_lib("nbody"), capacity=n)
        pop.set_option("pair_sym", sym)
        pop.set_tolerances(1e-8, 1e-8)
        pop.init(n, replay=False, seed=2)
        pop.rhs(0.0)
        pop.evolve_reset()
        t, h = 0.0, 0.1
        for _ in range(4):
            t, h = pop.evolve_apply("rkf45", t, 0.4, h)[:2]
        res.append((t, h, pop.evolve_stats(), pop.download()["y"].copy()))
        pop.close()
    assert res[0][2] == res[1][2]                                        # attempts, accepted, rejected
    assert abs(res[0][0] - res[1][0]) <= 1e-9 * res[0][0]                # the step size follows rmax continuously
    np.testing.assert_allclose(res[1][3], res[0][3], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("meth,n,sym", [("dde", 2500, 0), ("dde", 2500, 1), ("rkf45", 1100, 1), ("fe", 17, 1)])
def test_soft_repulsion_vs_oracle(model_lib, meth, n, sym):
    """a second antisymmetric coupling (models/repel2d.pop: softened repulsion in the plane, two of four variables coupled, a
    parameter inside the pair term, two %M mean fields next to it; restatement bit-identical to the reference binary,
    tests/test_oracle_port.py) through the ordered and the symmetric kernel against the CPU restatement over several tiles"""
    from odexp_b200 import capi
    from oracle import port
    opts = {"meth": meth, "h0": 0.05, "res": 3, "ps": n, "abstol": 1e-9, "reltol": 1e-9}
    pop = capi.Population(model_lib("repel2d"), capacity=n + 8)
    pop.set_option("pair_sym", sym)
    out = capi.odesolver(pop, (0, 0.5), meth=meth, res=3, popsize=n, h0=0.05, abstol=1e-9, reltol=1e-9, replay_ic=True)
    pop.close()
    ref = port.run("repel2d", opts, (0, 0.5))
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb) and out["report"]["steps"] == ref["report"]["steps"]
    np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, :4], tb[tb["step"] == 0]["cols"][:, :4])      # replayed ICs
    np.testing.assert_allclose(ta["cols"][:, :4], tb["cols"][:, :4], rtol=1e-10, atol=1e-12)
    scale = np.abs(tb["cols"][:, 4:6]).max()
    assert np.abs(ta["cols"][:, 4:6] - tb["cols"][:, 4:6]).max() <= 1e-11 * scale                             # psi columns
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-12)
