"""Pins for the clean-room GSL subset (oracle/minigsl.c) that do not need a libgsl: the Runge-Kutta order conditions of the
embedded pairs (first principles), convergence order of the whole evolve loop, and the Chebyshev routines against numpy.
Parity with a real GSL build stays unpinned; these tests pin the restatement to the published methods."""
import ctypes as C
import numpy as np
import pytest


def tableau(kind):
    from oracle import port
    L = port.lib()
    ah, b, c, ec = (np.zeros(5), np.zeros(25), np.zeros(6), np.zeros(6))
    P = C.POINTER(C.c_double)
    assert L.minigsl_tableau(kind, ah.ctypes.data_as(P), b.ctypes.data_as(P), c.ctypes.data_as(P), ec.ctypes.data_as(P)) == 0
    A = np.zeros((6, 6))
    A[1:, :5] = b.reshape(5, 5)
    nodes = np.concatenate([[0.0], ah])
    return A, nodes, c, ec


def order_conditions(A, c_nodes, w, order):
    """residuals of the rooted-tree order conditions up to `order` (<= 5) for weights w"""
    c = c_nodes
    Ac = A @ c
    res = [w.sum() - 1]
    if order >= 2: res += [w @ c - 1 / 2]
    if order >= 3: res += [w @ c**2 - 1 / 3, w @ Ac - 1 / 6]
    if order >= 4: res += [w @ c**3 - 1 / 4, w @ (c * Ac) - 1 / 8, w @ (A @ c**2) - 1 / 12, w @ (A @ Ac) - 1 / 24]
    if order >= 5:
        res += [w @ c**4 - 1 / 5, w @ (c**2 * Ac) - 1 / 10, w @ (Ac * Ac) - 1 / 20, w @ (c * (A @ c**2)) - 1 / 15,
                w @ (A @ c**3) - 1 / 20, w @ (c * (A @ Ac)) - 1 / 30, w @ (A @ (c * Ac)) - 1 / 40, w @ (A @ (A @ c**2)) - 1 / 60,
                w @ (A @ (A @ Ac)) - 1 / 120]
    return np.array(res)


@pytest.mark.parametrize("kind,name", [(2, "rkf45"), (3, "rkck")])
def test_embedded_pair_order_conditions(kind, name):
    A, nodes, c, ec = tableau(kind)
    np.testing.assert_allclose(A.sum(axis=1), nodes, atol=1e-15)                 # row-sum condition c_i = sum_j a_ij
    assert np.abs(order_conditions(A, nodes, c, 5)).max() < 1e-14                # propagated solution: order 5
    assert np.abs(order_conditions(A, nodes, c - ec, 4)).max() < 1e-14           # embedded solution c - ec: order 4
    assert abs(ec.sum()) < 1e-15                                                 # error weights sum to zero


def test_rk2_order_conditions_and_convergence():
    """GSL rk2 = embedded (2, 3): the propagated weights (1, 4, 1)/6 satisfy the conditions of order 3 on its three stages, the
    comparison solution c + ec = (0, 1, 0) (the midpoint rule; GSL's yerr = y_midpoint - y_new here) those of order 2, the error weights sum to zero; the whole evolve
    loop converges under tightening tolerances"""
    from oracle import port
    A, nodes, c, ec = tableau(0)
    A, nodes, c, ec = A[:3, :3], nodes[:3], c[:3], ec[:3]
    np.testing.assert_allclose(A.sum(axis=1), nodes, atol=1e-15)
    assert np.abs(order_conditions(A, nodes, c, 3)).max() < 1e-15
    assert np.abs(order_conditions(A, nodes, c + ec, 2)).max() < 1e-15
    np.testing.assert_allclose(c + ec, [0.0, 1.0, 0.0], atol=1e-16)
    assert abs(ec.sum()) < 1e-16
    ref = port.run("kuramoto", {"meth": "dde", "h0": 0.0005, "res": 2, "ps": 20}, (0, 2.0))["traj"]
    yref = ref[ref["step"] == 1]["cols"][:, 0]
    errs = []
    for tol in (1e-5, 1e-7, 1e-9):
        tr = port.run("kuramoto", {"meth": "rk2", "res": 2, "ps": 20, "abstol": tol, "reltol": tol, "h0": 0.5}, (0, 2.0))["traj"]
        errs.append(np.abs(tr[tr["step"] == 1]["cols"][:, 0] - yref).max())
    assert errs[2] < errs[1] < errs[0] and errs[0] < 1e-3 and errs[2] < 1e-7


def test_rk8pd_order_conditions_and_convergence():
    """Prince-Dormand RK8(7)13M as GSL's rk8pd: row sums equal the nodes, the rooted-tree conditions up to order 5 hold for both
    weight sets, the quadrature conditions sum_j w_j c_j^(q-1) = 1/q up to q = 8 for the propagated weights and q = 7 (not 8) for the
    comparison weights; the evolve loop converges at high order (a wrong digit in any coefficient breaks these at ~1e-9)"""
    from oracle import port
    L = port.lib()
    P = C.POINTER(C.c_double)
    ah, b, w8, w7 = np.zeros(12), np.zeros(144), np.zeros(13), np.zeros(13)
    assert L.minigsl_rk8pd_tableau(ah.ctypes.data_as(P), b.ctypes.data_as(P), w8.ctypes.data_as(P), w7.ctypes.data_as(P)) == 0
    A = np.zeros((13, 13)); A[1:, :12] = b.reshape(12, 12)
    nodes = np.concatenate([[0.0], ah])
    np.testing.assert_allclose(A.sum(axis=1), nodes, atol=4e-15)
    assert np.abs(order_conditions(A, nodes, w8, 5)).max() < 2e-15
    assert np.abs(order_conditions(A, nodes, w7, 5)).max() < 2e-15
    for q in range(1, 9):
        assert abs(w8 @ nodes**(q - 1) - 1 / q) < 1e-15
        if q < 8: assert abs(w7 @ nodes**(q - 1) - 1 / q) < 1e-15
    assert abs(w7 @ nodes**7 - 1 / 8) > 1e-5
    ref = port.run("kuramoto", {"meth": "dde", "h0": 0.00025, "res": 2, "ps": 20}, (0, 2.0))["traj"]
    yref = ref[ref["step"] == 1]["cols"][:, 0]
    errs, steps = [], []
    for tol in (1e-6, 1e-9, 1e-12):
        out = port.run("kuramoto", {"meth": "rk8pd", "res": 2, "ps": 20, "abstol": tol, "reltol": tol, "h0": 0.5}, (0, 2.0))
        tr = out["traj"]
        errs.append(np.abs(tr[tr["step"] == 1]["cols"][:, 0] - yref).max()); steps.append(out["report"]["steps"])
    assert errs[0] < 1e-6 and errs[1] < 1e-9 and errs[2] < 5e-12 and steps[2] < 40       # few steps even at 1e-12: high order


def test_evolve_convergence_and_control():
    """K = 0 turns the Kuramoto fixture into theta' = omega (integrated exactly by any consistent RK method); with coupling, tightening
    the tolerance by 100 must reduce the error against a tiny-step RK4 run by ~100 (5th-order pair under standard control)."""
    from oracle import port
    ref = port.run("kuramoto", {"meth": "dde", "h0": 0.0005, "res": 2, "ps": 20}, (0, 2.0))["traj"]
    yref = ref[ref["step"] == 1]["cols"][:, 0]
    errs = []
    for tol in (1e-6, 1e-8, 1e-10):
        out = port.run("kuramoto", {"meth": "rkf45", "res": 2, "ps": 20, "abstol": tol, "reltol": tol, "h0": 0.5}, (0, 2.0))
        tr = out["traj"]
        errs.append(np.abs(tr[tr["step"] == 1]["cols"][:, 0] - yref).max())
    assert errs[0] < 1e-5 and errs[1] < 1e-7 and errs[2] < 2e-9
    assert errs[2] < errs[1] < errs[0]


def test_cheb_against_numpy():
    """gsl_cheb_init / gsl_cheb_eval_n_err restated: coefficients equal numpy's Chebyshev interpolation at the same nodes"""
    from oracle import port
    L = port.lib()

    class GslFunction(C.Structure):
        _fields_ = [("function", C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)), ("params", C.c_void_p)]

    class ChebSeries(C.Structure):
        _fields_ = [("c", C.POINTER(C.c_double)), ("order", C.c_size_t), ("a", C.c_double), ("b", C.c_double),
                    ("order_sp", C.c_size_t), ("f", C.POINTER(C.c_double))]
    L.gsl_cheb_alloc.restype = C.POINTER(ChebSeries)
    L.gsl_cheb_alloc.argtypes = [C.c_size_t]
    L.gsl_cheb_init.argtypes = [C.POINTER(ChebSeries), C.POINTER(GslFunction), C.c_double, C.c_double]
    L.gsl_cheb_eval_n_err.argtypes = [C.POINTER(ChebSeries), C.c_size_t, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    order, a, b = 20, -3.0, 3.0
    cs = L.gsl_cheb_alloc(order)
    cb = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)(lambda x, p: float(np.sin(x)))
    F = GslFunction(cb, None)
    assert L.gsl_cheb_init(cs, C.byref(F), a, b) == 0
    coef = np.array([cs.contents.c[i] for i in range(order + 1)])
    cheb = np.polynomial.chebyshev.Chebyshev.interpolate(np.sin, order, domain=[a, b])
    expect = cheb.coef.copy(); expect[0] *= 2                                    # GSL convention: f = c0/2 + sum c_n T_n
    np.testing.assert_allclose(coef, expect, atol=1e-13)
    val, err = C.c_double(), C.c_double()
    for x in (-2.5, 0.3, 2.9):
        L.gsl_cheb_eval_n_err(cs, order, x, C.byref(val), C.byref(err))
        assert abs(val.value - np.sin(x)) < 1e-12 and err.value > 0
    L.gsl_cheb_free.argtypes = [C.POINTER(ChebSeries)]
    L.gsl_cheb_free(cs)
