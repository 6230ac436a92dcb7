"""GPU parity against the committed golden vectors (outputs of the UNCHANGED reference binary, see
tests/golden/make_golden.py).  Everything goes through the C-ABI (odx_odesolver via ctypes)."""
import ast
import numpy as np
import pytest
from conftest import load_golden

pytestmark = pytest.mark.gpu

# tolerance: north_star asks <= 1e-10 relative state error at t1 for non-chaotic models
RTOL_STATE = 1e-10
ATOL_MFD = 1e-12


def run_case(model_lib, name):
    from odexp_b200 import capi
    g = load_golden(name)
    meta = ast.literal_eval(str(g["meta"][0]))
    o = meta["opts"]
    lib = model_lib("kuramoto" if meta["fixture"] == "kuramoto_shipped" else meta["fixture"])
    pop = capi.Population(lib, capacity=max(1024, int(o["ps"])))
    kw = dict(meth=o["meth"], res=o["res"], popsize=o["ps"], h0=o.get("h0", 0.1), hmin=o.get("hmin", 1e-5),
              abstol=o.get("abstol", 1e-6), reltol=o.get("reltol", 1e-6))
    out = capi.odesolver(pop, meta["timespan"], replay_ic=True, **kw)
    pop.close()
    return g, out, meta


def check(g, out, rtol_state=RTOL_STATE, atol_mfd=ATOL_MFD):
    st = out["stats"]
    assert len(st) == len(g["stats_t"])
    np.testing.assert_array_equal(st["n"], g["stats_n"])
    np.testing.assert_array_equal(st["stop_flag"], g["stats_flag"])
    np.testing.assert_array_equal(st["event"], g["stats_event"])
    np.testing.assert_allclose(st["t"], g["stats_t"], rtol=1e-14, atol=1e-14)
    if g["stats_mfd"].shape[1]:
        np.testing.assert_allclose(st["mfd"], g["stats_mfd"], rtol=0, atol=atol_mfd)
    tr = out["traj"]
    assert tr["step"].max() + 1 == int(g["nframes"][0])
    first = tr[tr["step"] == 0]
    last = tr[tr["step"] == tr["step"].max()]
    np.testing.assert_array_equal(first["id"], g["first_id"])
    np.testing.assert_array_equal(last["id"], g["last_id"])
    # initial conditions are drawn through the restated libc stream: bit-identical
    nv = first["cols"].shape[1]
    assert nv == g["first_cols"].shape[1]
    return first, last


@pytest.mark.parametrize("name", ["kuramoto_dde_n64", "kuramoto_dde_n1000", "kuramoto_dde_h003_n257", "kuramoto_fe_n100"])
def test_kuramoto_fixed_step(model_lib, name):
    g, out, meta = run_case(model_lib, name)
    first, last = check(g, out)
    d = meta["dims"]
    # y and expr columns of the first frame are the ICs: must be bit-exact
    ycol, ecol = 0, d["nvar"] + d["naux"] + d["npsi"] + d["nmfd"]
    np.testing.assert_array_equal(first["cols"][:, ycol], g["first_cols"][:, ycol])
    np.testing.assert_array_equal(first["cols"][:, ecol], g["first_cols"][:, ecol])
    np.testing.assert_allclose(last["cols"], g["last_cols"], rtol=RTOL_STATE, atol=1e-13)


@pytest.mark.parametrize("name", ["kuramoto_rkck_n200", "kuramoto_rkf45_n200", "kuramoto_rk4_n50"])
def test_kuramoto_adaptive(model_lib, name):
    g, out, meta = run_case(model_lib, name)
    first, last = check(g, out, atol_mfd=1e-11)
    np.testing.assert_allclose(last["cols"], g["last_cols"], rtol=1e-9, atol=1e-12)


def test_nbody_dde(model_lib):
    g, out, meta = run_case(model_lib, "nbody_dde_n64")
    first, last = check(g, out)
    np.testing.assert_array_equal(first["cols"][:, :6], g["first_cols"][:, :6])
    # short horizon, chaotic N-body: stated tolerance 1e-8 relative on the state (pow(r2,1.5) -> r2*sqrt(r2))
    np.testing.assert_allclose(last["cols"], g["last_cols"], rtol=1e-8, atol=1e-12)


def test_nbody_rkf45(model_lib):
    g, out, meta = run_case(model_lib, "nbody_rkf45_n96")
    first, last = check(g, out)
    np.testing.assert_allclose(last["cols"], g["last_cols"], rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("meth,n,ts,h0", [("dde", 257, (0, 0.04), 0.02), ("dde", 1025, (0, 0.04), 0.02), ("dde", 4099, (0, 0.02), 0.02),
                                          ("rkf45", 257, (0, 0.05), 0.005), ("rkf45", 1025, (0, 0.05), 0.005), ("rkf45", 4099, (0, 0.002), 0.002)])
@pytest.mark.parametrize("sym", [0, 1])
def test_nbody_multi_tile_vs_oracle(model_lib, meth, n, ts, h0, sym):
    """the O(N^2) %C kernel beyond one tile: several i-tiles (256 particles) and j-segments (PAIR_TJ = 128), the self term's
    poison-and-redo in a tile other than the first, k_pair_finish adding the segments in list order -- against the CPU
    restatement (bit-identical to the reference binary for this fixture).  Stated tolerance for the chaotic N-body: 1e-8
    relative over this short horizon, identical accept / reject sequence for rkf45.  sym = 1 forces the symmetric kernel
    (k_pair_sym: every unordered pair once, the default from 16 tiles of 1024 particles on) through the same comparison."""
    from odexp_b200 import capi
    from oracle import port
    o = {"meth": meth, "h0": h0, "res": 2, "ps": n, "hmin": 0.0, "abstol": 1e-8, "reltol": 1e-8}
    pop = capi.Population(model_lib("nbody"), capacity=n + 64)
    pop.set_option("pair_sym", sym)
    out = capi.odesolver(pop, ts, meth=meth, h0=h0, res=2, popsize=n, hmin=0.0, abstol=1e-8, reltol=1e-8, replay_ic=True)
    pop.close()
    ref = port.run("nbody", o, ts)
    assert out["report"]["steps"] == ref["report"]["steps"]
    a, b = out["traj"], ref["traj"]
    np.testing.assert_array_equal(a["id"], b["id"])
    np.testing.assert_array_equal(a["cols"][a["step"] == 0][:, :6], b["cols"][b["step"] == 0][:, :6])      # replayed ICs: bit-identical
    np.testing.assert_allclose(a["cols"][:, :6], b["cols"][:, :6], rtol=1e-8, atol=1e-12)
    # the couplings themselves (columns 6..8 of a frame = psi): relative to the largest coupling of the frame
    scale = np.abs(b["cols"][:, 6:9]).max()
    assert np.abs(a["cols"][:, 6:9] - b["cols"][:, 6:9]).max() <= 1e-10 * scale


@pytest.mark.parametrize("n", [16387, 40001])
def test_nbody_psi_many_tiles_per_cta_vs_numpy(model_lib, n):
    """N large enough that every CTA of k_pair walks several j-tiles (more than 64 tiles of 128): the redo of the self term
    then happens in the middle of a CTA's j-loop, and the partial sums of 64 segments are added by k_pair_finish.  Checked for
    200 sampled particles against the numpy restatement of the generated %C walk (oracle/nbody_np.py)."""
    from odexp_b200 import capi
    from oracle import nbody_np
    pop = capi.Population(model_lib("nbody"), capacity=n)
    pop.init(n, replay=False, seed=5)
    pop.rhs(0.0)
    d = pop.download()
    pop.close()
    rng = np.random.default_rng(1)
    idx = np.unique(np.concatenate([[0, 1, 127, 128, 255, 256, n - 1, n - 2], rng.integers(0, n, 192)]))
    want = nbody_np.psi_subset(d["y"][:, :3], idx)
    scale = np.abs(want).max()
    assert np.abs(d["psi"][idx] - want).max() <= 1e-11 * scale


def test_config1_kuramoto_as_shipped(model_lib):
    """BASELINE config 1: the reference's examples/kuramoto.pop as shipped (N=1000, rkck, 801 output points over [0,160]);
    golden = the unchanged reference binary built from the shipped file.  160 time units of adaptive stepping: the GPU and
    CPU runs take the same accept/reject decisions, so the trajectories stay within 1e-8."""
    g, out, meta = run_case(model_lib, "kuramoto_shipped_rkck")
    st = out["stats"]
    assert len(st) == len(g["stats_t"])          # 1000 birth rows + one row per loop iteration (incl. the round-off sliver at t1)
    np.testing.assert_array_equal(st["stop_flag"], g["stats_flag"])
    np.testing.assert_allclose(st["t"], g["stats_t"], rtol=1e-13, atol=1e-12)
    np.testing.assert_allclose(st["mfd"], g["stats_mfd"], rtol=0, atol=1e-9)
    tr = out["traj"]
    last = tr[tr["step"] == tr["step"].max()]
    np.testing.assert_allclose(last["cols"], g["last_cols"], rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("name", ["ring_dde_n50", "ring_fe_n3"])
def test_ring_neighbours_ky(model_lib, name):
    """KY(name, degree): list neighbours on a ring (src/odexp.h:334-357).  The stages alternate between two buffers on the device
    because neighbours read the stage state while the next one is formed; golden = the unchanged reference binary."""
    g, out, meta = run_case(model_lib, name)
    first, last = check(g, out)
    np.testing.assert_array_equal(first["cols"][:, :2], g["first_cols"][:, :2])
    np.testing.assert_allclose(last["cols"], g["last_cols"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("n,meth", [(1, "dde"), (2, "dde"), (5, "fe"), (1025, "dde"), (4099, "fe")])
def test_ring_neighbours_vs_oracle(model_lib, n, meth):
    """ragged ring sizes (a ring of 1 or 2 particles wraps onto itself) against the CPU restatement"""
    from odexp_b200 import capi
    from oracle import port
    pop = capi.Population(model_lib("ring_diffusion"), capacity=max(1024, n + 8))
    out = capi.odesolver(pop, (0, 0.6), meth=meth, h0=0.02, res=4, popsize=n, replay_ic=True)
    ref = port.run("ring_diffusion", {"meth": meth, "h0": 0.02, "res": 4, "ps": n}, (0, 0.6))
    np.testing.assert_allclose(out["traj"]["cols"], ref["traj"]["cols"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-12)
    pop.close()


@pytest.mark.parametrize("meth", ["rk2", "rk8pd"])
@pytest.mark.parametrize("model,n,ts", [("kuramoto", 257, (0, 3.0)), ("nbody", 40, (0, 0.2)), ("ring_diffusion", 33, (0, 1.5))])
def test_gsl_rk2_rk8pd_vs_oracle(model_lib, model, n, ts, meth):
    """GSL's embedded rk2 (2, 3) and Prince-Dormand rk8pd (13 stages) through the general tableau kernel: same accept / reject
    sequence and states as the restatement (ring_diffusion: KY() neighbours, alternating stage buffers with an odd stage count)"""
    from odexp_b200 import capi
    from oracle import port
    tol = 1e-8 if model == "nbody" else 1e-6
    pop = capi.Population(model_lib(model), capacity=1024)
    out = capi.odesolver(pop, ts, meth=meth, res=4, popsize=n, abstol=tol, reltol=tol, hmin=0.0 if model == "nbody" else 1e-5,
                         h0=0.02 if model == "nbody" else 0.1, replay_ic=True)
    ref = port.run(model, {"meth": meth, "res": 4, "ps": n, "abstol": tol, "reltol": tol, "hmin": 0.0 if model == "nbody" else 1e-5,
                           "h0": 0.02 if model == "nbody" else 0.1}, ts)
    assert len(out["stats"]) == len(ref["stats"])
    assert out["report"]["steps"] == ref["report"]["steps"]
    np.testing.assert_allclose(out["traj"]["cols"], ref["traj"]["cols"], rtol=1e-8 if model == "nbody" else 1e-9, atol=1e-12)
    pop.close()


@pytest.mark.parametrize("meth,n", [("rkck", 33), ("rkf45", 200), ("rk4", 17)])
def test_ring_neighbours_adaptive_vs_oracle(model_lib, meth, n):
    """KY() under the embedded pairs and the step-doubling rk4 (same accept / reject sequence as the restatement, which is
    bit-identical to the reference binary for this fixture)"""
    from odexp_b200 import capi
    from oracle import port
    pop = capi.Population(model_lib("ring_diffusion"), capacity=1024)
    out = capi.odesolver(pop, (0, 1.5), meth=meth, res=4, popsize=n, replay_ic=True)
    ref = port.run("ring_diffusion", {"meth": meth, "res": 4, "ps": n}, (0, 1.5))
    assert len(out["stats"]) == len(ref["stats"])
    np.testing.assert_allclose(out["traj"]["cols"], ref["traj"]["cols"], rtol=1e-9, atol=1e-12)
    pop.close()


@pytest.mark.parametrize("replay", [True, False])
@pytest.mark.parametrize("meth,ts,res", [("rkf45", (0, 20.0), 11), ("rk8pd", (0, 20.0), 5), ("dde", (0, 2.0), 5)])
def test_const_arrays_indexed_by_id(model_lib, meth, ts, res, replay):
    """`decl const double a[] = {..}` arrays indexed by ID in the initial conditions and an {initial} parameter: the three bodies of
    the reference's shipped examples/Nbody.pop (src/odexp:660-674; restatement bit-identical to the reference binary).  ICs are
    set by the host replay path and by the device kernel; the orbit is followed over t = 20 (close encounters included)"""
    from odexp_b200 import capi
    from oracle import port
    opts = {"meth": meth, "h0": 0.05 if meth == "dde" else 0.1, "res": res, "ps": 3, "abstol": 1e-8, "reltol": 1e-8, "hmin": 0.0}
    pop = capi.Population(model_lib("nbody_const"), capacity=1024)
    out = capi.odesolver(pop, ts, meth=meth, res=res, popsize=3, abstol=1e-8, reltol=1e-8, hmin=0.0, h0=opts["h0"], replay_ic=replay)
    pop.close()
    ref = port.run("nbody_const", opts, ts)
    ta, tb = out["traj"], ref["traj"]
    np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, :6], tb[tb["step"] == 0]["cols"][:, :6])       # x_ic[ID] .. v0*vz0[ID]
    assert len(out["stats"]) == len(ref["stats"])
    assert out["report"]["steps"] == ref["report"]["steps"]
    np.testing.assert_allclose(ta["cols"], tb["cols"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("replay", [True, False])
@pytest.mark.parametrize("meth,n,res", [("dde", 20, 4), ("rkck", 33, 4), ("fe", 5, 3), ("dde", 4099, 3), ("rk8pd", 257, 3)])
def test_array_variables(model_lib, meth, n, res, replay):
    """`dX[i=0:NC]/dt`, `aux sq[i=0:NC]`, `expr w[i=0:NC]`, `init X[i=0:NC]` (one loop per entry, size_t iterator visible in the
    rhs: src/odexp:1331-1506) mixed with scalar entries, helper functions on the arrays (sum, dotprod), MA() in %M.  ICs from
    the replayed libc stream (bit-identical to the restatement) and from the device's keyed Philox draws."""
    from odexp_b200 import capi
    from oracle import port
    opts = {"meth": meth, "h0": 0.05, "res": res, "ps": n, "hmin": 1e-5}
    pop = capi.Population(model_lib("chain_array"), capacity=max(1024, n))
    out = capi.odesolver(pop, (0, 3), meth=meth, res=res, popsize=n, h0=0.05, hmin=1e-5, replay_ic=replay)
    pop.close()
    ref = port.run("chain_array", opts, (0, 3), rng_mode=0 if replay else 1)
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb) and len(out["stats"]) == len(ref["stats"])
    if replay:
        np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, :6], tb[tb["step"] == 0]["cols"][:, :6])      # z, X[0..3], u
        np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, -4:], tb[tb["step"] == 0]["cols"][:, -4:])    # w[0..3]
    assert out["report"]["steps"] == ref["report"]["steps"]
    np.testing.assert_allclose(ta["cols"], tb["cols"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("meth,n", [("dde", 10), ("rkf45", 21), ("fe", 1500), ("rk8pd", 300)])
def test_function_blocks(model_lib, meth, n):
    """`FUN name(args)` ... `end` blocks with statements, pointer arguments (`*x` passes an array) and a one-line function on an
    array (src/odexp:676-737, README FUN), called from AUX and from an array D/DT whose rhs also reads t and the iterator"""
    from odexp_b200 import capi
    from oracle import port
    # rk8pd at the default tolerance 1e-6 is not a parity case: its step-size control amplifies a 1e-10 relative change of the
    # tolerance into a 4e-8 change of the restatement's own solution (measured); at 1e-12 the step sequence no longer matters
    tol = 1e-12 if meth == "rk8pd" else 1e-6
    opts = {"meth": meth, "h0": 0.05, "res": 4, "ps": n, "hmin": 1e-5, "abstol": tol, "reltol": tol}
    pop = capi.Population(model_lib("fun_block"), capacity=max(1024, n))
    out = capi.odesolver(pop, (0, 3), meth=meth, res=4, popsize=n, h0=0.05, hmin=1e-5, abstol=tol, reltol=tol, replay_ic=True)
    pop.close()
    ref = port.run("fun_block", opts, (0, 3))
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb) and (meth == "rk8pd" or out["report"]["steps"] == ref["report"]["steps"])
    np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, :3], tb[tb["step"] == 0]["cols"][:, :3])
    # rk8pd: the restatement itself moves by 6e-12 under a 1e-8 relative change of the tolerance (states are O(1))
    np.testing.assert_allclose(ta["cols"], tb["cols"], rtol=1e-10, atol=2e-11 if meth == "rk8pd" else 1e-12)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=2e-11 if meth == "rk8pd" else 1e-12)


@pytest.mark.parametrize("meth,n,ts", [("dde", 30, (0, 2)), ("rkf45", 41, (0, 2)), ("fe", 7, (0, 0.5)), ("dde", 1300, (0, 0.5))])
def test_own_auxiliary_in_pairwise_coupling(model_lib, meth, n, ts):
    """`%C FRC MA(s)*OE(q)*sin(OY(x) - MY(x))`: the particle's own auxiliary value inside an all-pairs coupling is the one pass 1
    of pop_rhs left (zeroed mean fields -- the aux of the fixture reads a mean field -- and the previous psi); the factor is
    hoisted out of the N^2 sum by the emitter.  `%C DRG OA(w)*(OY(v) - MY(v))`: the OTHER particle's auxiliary value, evaluated
    per staged j-particle (w reads no mean field / coupling, so pass-1 and pass-2 values coincide).  Restatement bit-identical
    to the reference binary (tests/test_oracle_port.py); n = 1300 spans several j-tiles and i-tiles."""
    from odexp_b200 import capi
    from oracle import port
    h0 = 0.01 if meth == "fe" else 0.05
    opts = {"meth": meth, "h0": h0, "res": 3, "ps": n}
    pop = capi.Population(model_lib("pair_aux"), capacity=max(1024, n))
    out = capi.odesolver(pop, ts, meth=meth, res=3, popsize=n, h0=h0, replay_ic=True)
    pop.close()
    ref = port.run("pair_aux", opts, ts)
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb) and out["report"]["steps"] == ref["report"]["steps"]
    np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, :2], tb[tb["step"] == 0]["cols"][:, :2])
    np.testing.assert_allclose(ta["cols"], tb["cols"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name,rtol,atol", [("chain_array_dde_n20", 1e-10, 1e-13), ("fun_block_rkf45_n21", 1e-10, 1e-12),
                                            ("helper_calls_dde_n12", 1e-10, 1e-13), ("helper_calls_rkck_n9", 1e-9, 1e-12),
                                            ("delay_pop_dde_n8", 1e-10, 1e-13), ("delay_pop_dde_stop_n5", 1e-10, 1e-13),
                                            ("pair_aux_dde_n30", 1e-10, 1e-12), ("nbody_const_rkf45", 1e-8, 1e-10),
                                            ("repel2d_dde_n50", 1e-10, 1e-12), ("repel2d_rkf45_n33", 1e-9, 1e-12),
                                            ("interp_drive_dde_n12", 1e-10, 1e-13),
                                            ("lowrank_aux_dde_n100", 1e-10, 1e-9), ("lrexp_dde_n200", 1e-10, 1e-9)])
def test_model_language_fixtures_against_reference_outputs(model_lib, name, rtol, atol):
    """the model-language fixtures of DESIGN 1.1 against what the UNCHANGED reference binary wrote for them (tests/golden): stats
    rows, ids, bit-identical initial state, final frame, current.plot.  atol 1e-9 for the low-rank cases covers the psi column
    (values ~ 1e2, expansion round-off, see test_gpu_lowrank.py); states are checked to rtol"""
    g, out, meta = run_case(model_lib, name)
    first, last = check(g, out, atol_mfd=1e-11)
    d = meta["dims"]
    np.testing.assert_array_equal(first["cols"][:, :d["nvar"]], g["first_cols"][:, :d["nvar"]])
    np.testing.assert_allclose(last["cols"][:, :d["nvar"]], g["last_cols"][:, :d["nvar"]], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(last["cols"], g["last_cols"], rtol=rtol, atol=atol)
    np.testing.assert_allclose(out["quick"], g["quick"], rtol=rtol, atol=atol, equal_nan=True)


@pytest.mark.parametrize("meth,n", [("dde", 10), ("rkf45", 23), ("fe", 3000)])
def test_file_table(model_lib, meth, n):
    """`FILE tab table_drive.csv {nrow, ncol, sep, skip}` (declare_array_from_file, src/odexp:581-641): the table becomes a constexpr
    constant of the generated functions, indexed by the particle ID in AUX, D/DT and the initial conditions"""
    from odexp_b200 import capi
    from oracle import port
    opts = {"meth": meth, "h0": 0.05, "res": 4, "ps": n}
    pop = capi.Population(model_lib("table_drive"), capacity=max(1024, n))
    out = capi.odesolver(pop, (0, 3), meth=meth, res=4, popsize=n, h0=0.05, replay_ic=True)
    pop.close()
    ref = port.run("table_drive", opts, (0, 3))
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb) and out["report"]["steps"] == ref["report"]["steps"]
    np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, 0], tb[tb["step"] == 0]["cols"][:, 0])
    np.testing.assert_array_equal(ta["cols"][:, 1], tb["cols"][:, 1])                  # aux amp = tab[ID % 4][1]
    np.testing.assert_allclose(ta["cols"], tb["cols"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("meth,n", [("dde", 12), ("rkck", 31), ("fe", 2500)])
def test_lag_linear_chain(model_lib, meth, n):
    """`LAG xd {root = x, len = L, mean = tau, init = 0.5}` (src/odexp:383-391): rewritten into an array of hidden dynamical
    variables driven by linchaindelay() (src/utils_odexp.c:162-167), the auxiliary variable at the end of the chain and their
    initial conditions; restatement bit-identical to the reference binary"""
    from odexp_b200 import capi
    from oracle import port
    opts = {"meth": meth, "h0": 0.05, "res": 5, "ps": n}
    pop = capi.Population(model_lib("lag_chain"), capacity=max(1024, n))
    out = capi.odesolver(pop, (0, 4), meth=meth, res=5, popsize=n, h0=0.05, replay_ic=True)
    pop.close()
    ref = port.run("lag_chain", opts, (0, 4))
    ta, tb = out["traj"], ref["traj"]
    assert len(ta) == len(tb) and out["report"]["steps"] == ref["report"]["steps"]
    np.testing.assert_array_equal(ta[ta["step"] == 0]["cols"][:, :6], tb[tb["step"] == 0]["cols"][:, :6])
    np.testing.assert_allclose(ta["cols"], tb["cols"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("meth,n,ts,h0", [("dde", 16, (0, 3.1), 0.03), ("dde", 1, (0, 2.0), 0.05), ("dde", 5, (0, 1.0, 2.5), 0.05), ("fe", 7, (0, 1.0), 0.02),
                                          ("rkck", 7, (0, 1.0), 0.05)])
def test_delay_history_stacks_vs_oracle(model_lib, meth, n, ts, h0):
    """delay(t, x, tau, ic, &hist[ID]) (src/utils_odexp.c:205-240): one circular history of 1000 samples per particle on the device,
    filled from the history function at t0 and while an event flag is pending (an interior TIMESPAN value), pushed twice per first
    evaluation of a dde step, interpolated with the reference's index arithmetic; the other steppers only look up (as in the
    reference, where only dde_apply sets DELAY_FIRST_EVAL).  Against the restatement, which is bit-identical to the reference binary."""
    from odexp_b200 import capi
    from oracle import port
    opts = {"meth": meth, "h0": h0, "res": 7, "ps": n}
    pop = capi.Population(model_lib("delay_pop"), capacity=64)
    out = capi.odesolver(pop, ts, meth=meth, h0=h0, res=7, popsize=n, replay_ic=True)
    pop.close()
    ref = port.run("delay_pop", opts, ts)
    assert len(out["stats"]) == len(ref["stats"]) and out["report"]["steps"] == ref["report"]["steps"]
    np.testing.assert_array_equal(out["stats"]["stop_flag"], ref["stats"]["stop_flag"])
    np.testing.assert_allclose(out["traj"]["cols"], ref["traj"]["cols"], rtol=1e-10 if meth != "rkck" else 1e-9, atol=1e-13)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("meth,tol,bound", [("rkck", 1e-6, 2e-4), ("rkck", 1e-10, 5e-8), ("rkf45", 1e-8, 5e-5)])
def test_interp_forcing_adaptive(model_lib, meth, tol, bound):
    """interp() forcing with kinks under the adaptive steppers.  On such a right-hand side GSL's step control is chaotic in the
    last bits: the restatement itself takes 191 / 201 / 209 steps at tolerance 1e-8 when h0 is changed by 0, 1e-15, 1e-14 relative
    (a stage point moving across a node changes the error estimate by O(1)), and its result moves by 3e-7.  GPU and CPU therefore
    agree to the accuracy the tolerance buys (observed 10-30 x tol), not to round-off; the fixed-step cases of the same fixture
    (dde golden above, oracle fe / dde) are exact to 1e-10."""
    from odexp_b200 import capi
    from oracle import port
    o = {"meth": meth, "h0": 0.05, "res": 5, "ps": 7, "hmin": 1e-5, "abstol": tol, "reltol": tol}
    pop = capi.Population(model_lib("interp_drive"), capacity=1024)
    out = capi.odesolver(pop, (0, 5), meth=meth, res=5, popsize=7, h0=0.05, hmin=1e-5, abstol=tol, reltol=tol, replay_ic=True)
    pop.close()
    ref = port.run("interp_drive", o, (0, 5))
    a, b = out["traj"], ref["traj"]
    np.testing.assert_array_equal(a["id"], b["id"])
    np.testing.assert_array_equal(a[a["step"] == 0]["cols"][:, 0], b[b["step"] == 0]["cols"][:, 0])
    assert np.abs(a["cols"][:, 0] - b["cols"][:, 0]).max() <= bound
    assert abs(out["report"]["steps"] - ref["report"]["steps"]) <= 0.15 * ref["report"]["steps"]


def test_hmin_warning_is_reported(model_lib):
    """`h < hmin` after an adaptive step: the run goes on with h = hmin and the driver reports the first time it happened, like the
    reference's one-time "warning: hmin reached at t = ..." (src/methods_odexp.c:604-612); the drop-in branch prints it from the report"""
    from odexp_b200 import capi
    got = {}
    for hmin in (1e-5, 2e-2):
        pop = capi.Population(model_lib("interp_drive"), capacity=1024)
        out = capi.odesolver(pop, (0, 5), meth="rkck", res=5, popsize=7, h0=0.05, hmin=hmin, abstol=1e-8, reltol=1e-8, replay_ic=True)
        pop.close()
        got[hmin] = out["report"]
    assert got[1e-5]["hmin_reached"] == 0
    assert got[2e-2]["hmin_reached"] == 1 and 0.0 < got[2e-2]["hmin_t"] < 5.0
