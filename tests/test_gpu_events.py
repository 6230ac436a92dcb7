"""GPU parity for the stochastic paths: Euler-Maruyama (DWDT) and birth / death (SSA, tau-leaping), through the C-ABI,
against the CPU restatement run with the SAME keyed Philox draws (oracle rng_mode 1; the restatement itself is pinned
bit-for-bit to the unchanged reference with the serial libc stream, tests/test_oracle_port.py).

Required: identical event logs (N, stop flag, parent id, +-1, child id per stats row) and identical particle ids in
every trajectory frame; states within 1e-10 relative (libm vs device log/sqrt/cos in Box-Muller)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_both(model_lib, model, opts, ts, capacity=None):
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib(model)
    pop = capi.Population(lib, capacity=capacity or max(4096, 4 * opts["ps"]))
    out = capi.odesolver(pop, ts, meth=opts["meth"], res=opts["res"], popsize=opts["ps"], h0=opts.get("h0", 0.1),
                         hmin=opts.get("hmin", 1e-5), ssahmin=opts.get("ssahmin", 0.0), aleap=opts.get("aleap", 0.01),
                         seed=opts.get("seed", 3141592), replay_ic=False)
    pop.close()
    ref = port.run(model, opts, ts, rng_mode=1)
    return out, ref


def compare(out, ref, rtol=1e-10, atol=1e-12):
    sa, sb, ta, tb = out["stats"], ref["stats"], out["traj"], ref["traj"]
    assert len(sa) == len(sb), (len(sa), len(sb))
    np.testing.assert_array_equal(sa["n"], sb["n"])
    np.testing.assert_array_equal(sa["stop_flag"], sb["stop_flag"])
    np.testing.assert_array_equal(sa["event"], sb["event"])
    np.testing.assert_allclose(sa["t"], sb["t"], rtol=1e-12, atol=1e-13)
    assert len(ta) == len(tb)
    np.testing.assert_array_equal(ta["id"], tb["id"])
    np.testing.assert_array_equal(ta["step"], tb["step"])
    np.testing.assert_allclose(ta["cols"], tb["cols"], rtol=rtol, atol=atol)


def test_kuramoto_device_initial_conditions(model_lib):
    """ode_init_conditions on the device (keyed Philox randN / rand01) == restatement with the same keys"""
    out, ref = run_both(model_lib, "kuramoto", {"meth": "dde", "h0": 0.1, "res": 3, "ps": 777}, (0, 0.4))
    compare(out, ref, rtol=1e-11, atol=1e-13)


def test_sde_euler_maruyama_no_events(model_lib):
    out, ref = run_both(model_lib, "vdp_sde", {"meth": "fe", "h0": 0.01, "res": 6, "ps": 300, "ssahmin": 0.0, "seed": 11}, (0, 0.05))
    compare(out, ref)


@pytest.mark.parametrize("ps,seed", [(100, 3141592), (1000, 5), (1, 7), (3000, 9)])
def test_tau_leaping(model_lib, ps, seed):
    opts = {"meth": "fe", "h0": 0.01, "res": 41, "ps": ps, "ssahmin": 1.0, "seed": seed}
    out, ref = run_both(model_lib, "vdp_sde", opts, (0, 4))
    ev = out["stats"]["event"][:, 1]
    assert (ev == 1).sum() > ps                               # births / replications beyond the initial rows
    if ps >= 100:
        assert (ev == -1).sum() > 0                           # and deaths
    compare(out, ref)


@pytest.mark.parametrize("ps,seed", [(50, 3141592), (200, 21)])
def test_ssa_single_events(model_lib, ps, seed):
    opts = {"meth": "fe", "h0": 0.01, "res": 11, "ps": ps, "ssahmin": 0.0, "seed": seed}
    out, ref = run_both(model_lib, "vdp_sde", opts, (0, 2))
    compare(out, ref)


@pytest.mark.parametrize("meth,opts,ts", [
    ("fe", {"h0": 0.01, "res": 21, "ps": 10, "ssahmin": 0.0}, (0, 2)),                  # SSA single events
    ("fe", {"h0": 0.01, "res": 11, "ps": 40, "ssahmin": 1.0}, (0, 2)),                  # tau-leaping
    ("dde", {"h0": 0.05, "res": 9, "ps": 500, "ssahmin": 1.0, "seed": 4}, (0, 2)),
    ("rkck", {"h0": 0.1, "res": 5, "ps": 12, "ssahmin": 1.0}, (0, 1)),
    ("rk8pd", {"h0": 0.1, "res": 5, "ps": 300, "ssahmin": 1.0, "seed": 8}, (0, 1)),
])
def test_vect_meanfields_with_birth_death(model_lib, meth, opts, ts):
    """%M ... {vect} (src/odexp:886-889): POP_SIZE as a mean field and a vect field reading a looped one, while the
    population size changes under SSA / tau-leaping; the mean-field columns of every trajectory row are compared too"""
    opts = dict(opts, meth=meth)
    out, ref = run_both(model_lib, "vect_meanfield", opts, ts)
    assert len(np.unique(out["stats"]["n"])) > 1, "population size must change during the run"
    compare(out, ref)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=1e-11, atol=1e-13)
    solver_rows = out["stats"]["event"][:, 1] == 0          # event rows are written before the mean fields are re-evaluated
    np.testing.assert_array_equal(out["stats"]["mfd"][solver_rows, 1], out["stats"]["n"][solver_rows].astype(float))      # TOTAL == POP_SIZE


def test_population_can_die_out(model_lib):
    """edge: every particle dies (huge death rate, no births): the loop stops with an empty population"""
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib("vdp_sde")
    mu = [1.0, 0.1, 50.0, 0.0]
    opts = {"meth": "fe", "h0": 0.01, "res": 11, "ps": 64, "ssahmin": 1.0, "aleap": 1.0}
    pop = capi.Population(lib, capacity=1024)
    out = capi.odesolver(pop, (0, 2), meth="fe", res=11, popsize=64, h0=0.01, ssahmin=1.0, aleap=1.0, replay_ic=False, mu=mu)
    ref = port.run("vdp_sde", opts, (0, 2), rng_mode=1, mu=mu)
    pop.close()
    np.testing.assert_array_equal(out["stats"]["n"], ref["stats"]["n"])
    np.testing.assert_array_equal(out["stats"]["event"], ref["stats"]["event"])


def test_exclusive_scan_and_compaction_large(model_lib):
    """size-independent property at a population larger than one scan tile grid: ids stay sorted (stable compaction
    preserves list order), are unique, and N matches the event log arithmetic"""
    from odexp_b200 import capi
    lib = model_lib("vdp_sde")
    n = 1 << 18
    pop = capi.Population(lib, capacity=2 * n)
    pop.set_seed(5)
    pop.init(n, replay=False, seed=5)
    pop.rhs(0.0)
    rows = pop.tauleap_apply(0.0, 2.0, 0)
    n_local, _, max_id = pop.size()
    d = pop.download()
    ids = d["id"]
    assert np.all(np.diff(ids) > 0) and len(ids) == n_local
    deaths = sum(1 for r in rows if r[2][1] == -1)
    births = sum(1 for r in rows if r[2][1] == 1)
    assert n_local == n - deaths + births and rows[-1][0] == n_local
    assert max_id == n + births
    pop.close()


def test_iteration_method(model_lib):
    """meth iteration (discrete map population, reference iteration_apply incl. its in-place two-pass behaviour)"""
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib("logistic_map")
    pop = capi.Population(lib, capacity=1024)
    out = capi.odesolver(pop, (0, 20), meth="iteration", popsize=333, replay_ic=True)
    pop.close()
    ref = port.run("logistic_map", {"meth": "iteration", "ps": 333}, (0, 20))
    assert len(out["stats"]) == len(ref["stats"]) and len(out["traj"]) == len(ref["traj"])
    np.testing.assert_allclose(out["stats"]["t"], ref["stats"]["t"], rtol=0, atol=0)
    # the logistic map at r = 3.7 is chaotic (Lyapunov exponent ~0.35 per iteration): round-off grows e^(0.35 n)
    first = out["traj"]["step"] <= 5
    np.testing.assert_allclose(out["traj"]["cols"][first], ref["traj"]["cols"][first], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(out["traj"]["cols"], ref["traj"]["cols"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("meth,opts,ts,tol", [
    ("dde", dict(h0=0.05, res=25, popsize=20), (0, 1.5, 3.25, 6), 1e-12),
    ("rkck", dict(h0=0.1, res=13, popsize=7), (0, 0.5, 4.0, 4.1, 7), 1e-9),
    ("fe", dict(h0=0.01, res=9, popsize=1000), (0, 2.5, 1.0, 4), 1e-11),          # unsorted stops
])
@pytest.mark.parametrize("replay", [True, False])
def test_stopping_times_stopif_ist0(model_lib, meth, opts, ts, tol, replay):
    """interior TIMESPAN values (STOP_TSPAN) and STOPIF (STOP_COND) cut the interval and re-initialise the living particles,
    IST0 selects the afterwards arm (src/methods_odexp.c:222-236, 453-459, 704-720).  replay: the serial libc stream of the
    reference; keyed: the same Philox draws in the restatement (rng_mode 1)."""
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib("reset_decay")
    assert lib.info.has_stopif == 1
    pop = capi.Population(lib, capacity=2048)
    out = capi.odesolver(pop, ts, meth=meth, replay_ic=replay, **opts)
    pop.close()
    o = {"meth": meth, "h0": opts["h0"], "res": opts["res"], "ps": opts["popsize"]}
    ref = port.run("reset_decay", o, ts, rng_mode=0 if replay else 1)
    assert len(out["stats"]) == len(ref["stats"]) and len(out["traj"]) == len(ref["traj"])
    np.testing.assert_array_equal(out["stats"]["stop_flag"], ref["stats"]["stop_flag"])
    assert (ref["stats"]["stop_flag"] & 8).any() and (ref["stats"]["stop_flag"] & 4).any()
    np.testing.assert_allclose(out["stats"]["t"], ref["stats"]["t"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=tol, atol=tol)
    np.testing.assert_array_equal(out["traj"]["id"], ref["traj"]["id"])
    np.testing.assert_allclose(out["traj"]["cols"], ref["traj"]["cols"], rtol=tol, atol=tol)


def test_stopping_time_keeps_evolve_state(model_lib):
    """the reference does not re-allocate its gsl evolve object at a stopping time: the first embedded step after the reset
    starts from the derivative of the state before the reset (dydt_in <- dydt_out); same here"""
    from odexp_b200 import capi
    from oracle import port
    lib = model_lib("kuramoto")
    pop = capi.Population(lib, capacity=256)
    out = capi.odesolver(pop, (0, 0.7, 2.0), meth="rkf45", h0=0.1, res=6, popsize=30, replay_ic=True)
    pop.close()
    ref = port.run("kuramoto", {"meth": "rkf45", "h0": 0.1, "res": 6, "ps": 30}, (0, 0.7, 2.0))
    assert len(out["stats"]) == len(ref["stats"]) and len(out["traj"]) == len(ref["traj"])
    np.testing.assert_array_equal(out["stats"]["stop_flag"], ref["stats"]["stop_flag"])
    np.testing.assert_allclose(out["stats"]["t"], ref["stats"]["t"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["traj"]["cols"], ref["traj"]["cols"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("model,opts,ts,axes,particle", [
    ("kuramoto", {"meth": "dde", "h0": 0.05, "res": 6, "ps": 20}, (0, 1.0), (-1, 0, 1), 0),                 # reference defaults T, [0], [1]
    ("kuramoto", {"meth": "dde", "h0": 0.05, "res": 6, "ps": 20}, (0, 1.0), (2, 4, 5), 0),                  # an auxiliary, two mean fields
    ("kuramoto", {"meth": "rkck", "h0": 0.1, "res": 6, "ps": 20}, (0, 1.0), (0, 1, 3), 3),                  # y[cp * tx + ng]: another particle's variable
    ("kuramoto", {"meth": "dde", "h0": 0.05, "res": 4, "ps": 5}, (0, 1.0), (0, -1, 2), 4),                  # past the end of y: NaN
    ("nbody", {"meth": "dde", "h0": 0.05, "res": 3, "ps": 12}, (0, 0.1), (7, 2, 8), 5),                     # couplings of particle 5
    ("vdp_sde", {"meth": "fe", "h0": 0.01, "res": 41, "ps": 30, "ssahmin": 1.0}, (0, 8), (1, 0, -1), 21),   # the plotted particle dies
    ("vdp_sde", {"meth": "fe", "h0": 0.01, "res": 21, "ps": 10, "ssahmin": 0.0}, (0, 4), (-1, 1, 0), 11),   # ... or is born later
    ("reset_decay", {"meth": "dde", "h0": 0.05, "res": 25, "ps": 20}, (0, 1.5, 3.25, 6), (-1, 0, 2), 1),    # rows after stopping times
])
def test_quick_buffer(model_lib, model, opts, ts, axes, particle):
    """current.plot, the three-column buffer the reference's gnuplot UI plots (fwrite_quick, src/methods_odexp.c:1353-1436):
    one row at t0, after every output interval, after every birth / death sweep and after every stop; columns by `opt x|y|z`
    for the particle `opt particle` (looked up by id on the device).  The restatement is pinned to the reference binary
    (tests/test_oracle_port.py::test_quick_buffer_matches_reference_binary)."""
    from odexp_b200 import capi
    from oracle import port
    keyed = model == "vdp_sde"
    pop = capi.Population(model_lib(model), capacity=4096)
    out = capi.odesolver(pop, ts, meth=opts["meth"], res=opts["res"], popsize=opts["ps"], h0=opts["h0"], ssahmin=opts.get("ssahmin", 0.0),
                         replay_ic=not keyed, plot=axes, particle=particle)
    if model == "kuramoto":
        assert pop.find_id(particle) == particle and pop.find_id(10 ** 6) == -1
    pop.close()
    ref = port.run(model, opts, ts, rng_mode=1 if keyed else 0, plot=axes, particle=particle)
    qa, qb = out["quick"], ref["quick"]
    assert qa.shape == qb.shape == (len(out["traj"]["step"]) and out["traj"]["step"].max() + 1, 3)
    np.testing.assert_array_equal(np.isnan(qa), np.isnan(qb))
    np.testing.assert_allclose(qa, qb, rtol=1e-10, atol=1e-12, equal_nan=True)
    if model == "vdp_sde":
        assert np.isnan(qb).any() and not np.isnan(qb).all()


@pytest.mark.parametrize("meth,ssahmin,ps,seed", [("fe", 1.0, 20, 3141592), ("fe", 1.0, 2000, 17), ("fe", 0.0, 20, 5), ("rkck", 1.0, 50, 9)])
def test_lineage_sister_macros(model_lib, meth, ssahmin, ps, seed):
    """SE / SA / SY and ISDAUGHTER in the expressions and initial conditions of replicating particles (the mother is re-initialised
    first with no sister, the daughter then reads the mother's NEW expressions and state and her OLD auxiliaries:
    src/methods_odexp.c:1697-1717, src/odexp.h:296-326), exprand / poisson draws inside expressions; same keyed draws on both sides"""
    opts = {"meth": meth, "h0": 0.01 if meth == "fe" else 0.1, "res": 31 if meth == "fe" else 13, "ps": ps, "ssahmin": ssahmin, "seed": seed}
    out, ref = run_both(model_lib, "lineage", opts, (0, 6))
    ev = ref["stats"]["event"]
    assert ((ev[:, 1] == 1) & (ev[:, 0] >= 0)).sum() > 0, "case must contain replications"
    gen = ref["traj"]["cols"][:, -2]
    assert gen.max() >= 2, "case must contain second-generation daughters"
    compare(out, ref)
    np.testing.assert_allclose(out["stats"]["mfd"], ref["stats"]["mfd"], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("n", [1, 7, 2047, 2048, 2049, 100003])
def test_ssa_pick_device_equals_host(model_lib, n):
    """the SSA event pick on the device (tile sums + one search block, 16 bytes back) chooses the same event as the host form
    that restates ncumsum + `while (u > r[k]) k++` in the reference's order (src/methods_odexp.c:1596, 1764-1777), over uniforms
    that hit the death range, the replication range, the birth entry and both ends"""
    from odexp_b200 import capi
    lib = model_lib("vdp_sde")
    pops = []
    for dev in (0, 1):
        p = capi.Population(lib, capacity=n + 256)
        p.set_option("ssa_device", dev)
        p.init(n, replay=False, seed=21)
        t = p.run_fixed("fe", 0.0, 0.01, 30)            # spread x around 0 so that the replication rates b*(x > 0) differ
        p.rhs(t)
        pops.append(p)
    us = [0.0, 1e-300, 1e-9, 0.013, 0.25, 0.49999, 0.5, 0.731, 0.9, 0.99, 0.999999, 1.0 - 1e-16, 1.0]
    rng = np.random.default_rng(3)
    us += list(rng.random(12))
    for u in us:
        ra = pops[0].ssa_apply(t, u)
        rb = pops[1].ssa_apply(t, u)
        assert ra == rb, (u, ra, rb)
        assert pops[0].size() == pops[1].size()
        for p in pops:
            p.rhs(t)
    da, db = pops[0].download(), pops[1].download()
    np.testing.assert_array_equal(da["id"], db["id"])
    np.testing.assert_array_equal(da["y"], db["y"])
    for p in pops:
        p.close()


@pytest.mark.parametrize("model,n,dt,sweeps", [("vdp_sde", 5, 8.0, 4), ("vdp_sde", 70001, 2.0, 4), ("vdp_sde", 1 << 20, 1.0, 3),
                                               ("vdp_sde", 300000, 40.0, 2), ("lineage", 3001, 4.0, 4)])
def test_fused_sweep_equals_per_generation_sweep(model_lib, model, n, dt, sweeps):
    """the tau-leaping sweep as one cooperative launch (k_leap_sweep: device-side generation loop, chunk totals + grid barrier
    instead of the scan kernels, rows written into pinned host memory) against the per-generation kernels, which the oracle /
    reference-binary parity tests pinned first (src/methods_odexp.c:1662-1762): identical event rows, ids, states and max_id
    after every sweep.  dt = 40 makes nearly every particle an event: more rows than the pinned buffer holds at first (the rows
    are then rebuilt by the scan kernels); lineage: SY / SE / SA and ISDAUGHTER in the daughters' initial conditions."""
    from odexp_b200 import capi
    res = []
    for fused in (0, 1):
        pop = capi.Population(model_lib(model), capacity=3 * n + 4096)
        pop.set_option("leap_fused", fused)
        pop.init(n, replay=False, seed=9)
        t = pop.run_fixed("fe", 0.0, 0.01, 5)
        got = []
        for k in range(sweeps):
            pop.rhs(t)
            rows = pop.tauleap_apply(t, dt, k, as_array=True)
            d = pop.download()
            got.append((rows.tobytes(), d["id"].copy(), d["y"].copy(), d["expr"].copy(), pop.size()))
        res.append(got)
        pop.close()
    for a, b in zip(*res):
        assert a[4] == b[4]
        assert a[0] == b[0]
        np.testing.assert_array_equal(a[1], b[1])
        np.testing.assert_array_equal(a[2], b[2])
        np.testing.assert_array_equal(a[3], b[3])
    assert len(res[1][0][0]) > 0 or n < 10


def test_fused_sweep_capacity_error(model_lib):
    """daughters beyond the capacity given to odx_create: an error from the call, not a write past the buffers"""
    from odexp_b200 import capi
    n = 50000
    lib = model_lib("vdp_sde")
    pop = capi.Population(lib, capacity=n + 32)
    mu = list(lib.par_defaults)                 # par mu, sigma, d, b (models/vdp_sde.pop): no deaths, replication rate 0.5 where x > 0
    mu[2], mu[3] = 0.0, 0.5
    pop.set_params(mu)
    pop.init(n, replay=False, seed=9)
    pop.rhs(0.0)
    with pytest.raises(capi.OdxError, match="capacity"):
        pop.tauleap_apply(0.0, 1.0, 0, as_array=True)
    pop.close()
