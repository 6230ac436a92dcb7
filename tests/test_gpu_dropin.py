"""The drop-in boundary end to end: the reference's OWN model.out, linked against the reference sources with the B200 branch
added to odesolver() (integration/apply_gpu_branch.py + integration/odexp_gpu_branch.c; built into oracle/_ref/<fixture>/
model_gpu.out by oracle/build_ref.py), run with ODEXP_CUDA_MODEL pointing at the model library -- against the same binary's
unchanged CPU path.  What the UI reads afterwards must be the same: .odexp/stats.dat, traj.dat, current.plot and SIM->pop."""
import os
import numpy as np
import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _both(model_lib, fixture, opts, ts):
    from oracle import refrun, port
    if not refrun.have_dropin(fixture):
        pytest.skip("oracle/_ref/%s/model_gpu.out not built (needs /root/reference at build time)" % fixture)
    lib = model_lib(fixture)
    dims = port.DIMS[fixture]
    cpu = refrun.run(fixture, opts, ts, dims=dims, exe="model_gpu.out")                    # no ODEXP_CUDA_MODEL: the unchanged CPU path
    gpu = refrun.run(fixture, opts, ts, dims=dims, exe="model_gpu.out",
                     extra_env={"ODEXP_CUDA_MODEL": lib.path, "ODEXP_CUDA_DUMP_POP": "pop_dump.bin"})
    assert "on the GPU" in gpu["stdout"], gpu["stdout"][-2000:]
    assert "on the GPU" not in cpu["stdout"]
    return cpu, gpu, dims


def _check_files(cpu, gpu, dims, rtol, atol_mfd=1e-12):
    a, b = cpu["stats"], gpu["stats"]
    assert len(a) == len(b)
    for k in ("n", "stop_flag", "event"):
        np.testing.assert_array_equal(a[k], b[k])                                          # integer columns: byte-identical
    np.testing.assert_allclose(b["t"], a["t"], rtol=1e-14, atol=1e-14)
    if dims["nmfd"]:
        np.testing.assert_allclose(b["mfd"], a["mfd"], rtol=0, atol=atol_mfd)
    ta, tb = cpu["traj"], gpu["traj"]
    assert len(ta) == len(tb)
    np.testing.assert_array_equal(ta["step"], tb["step"])
    np.testing.assert_array_equal(ta["id"], tb["id"])
    np.testing.assert_allclose(tb["t"], ta["t"], rtol=1e-14, atol=1e-14)
    f0a, f0b = ta[ta["step"] == 0]["cols"], tb[tb["step"] == 0]["cols"]
    nv, ne = dims["nvar"], dims["nexpr"]
    assert f0a[:, :nv].tobytes() == f0b[:, :nv].tobytes()                                   # replayed initial conditions: bit-identical
    if ne:
        assert f0a[:, -ne:].tobytes() == f0b[:, -ne:].tobytes()
    np.testing.assert_allclose(tb["cols"], ta["cols"], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(gpu["quick"], cpu["quick"], rtol=rtol, atol=1e-12, equal_nan=True)  # current.plot


def _check_mirror(gpu, dims):
    """SIM->pop after the GPU run (walked by the branch with the reference's list, ODEXP_CUDA_DUMP_POP) = the last traj.dat frame"""
    nv, ne, na, npsi, nm = dims["nvar"], dims["nexpr"], dims["naux"], dims["npsi"], dims["nmfd"]
    rec = np.dtype([("id", "<i4"), ("y", "<f8", (nv,)), ("expr", "<f8", (ne,)), ("aux", "<f8", (na,)), ("psi", "<f8", (npsi,)),
                    ("death", "<f8"), ("repli", "<f8")])
    raw = gpu["pop_dump"]
    pop = np.frombuffer(raw[:-4], dtype=rec)
    max_id = int(np.frombuffer(raw[-4:], dtype="<i4")[0])
    tr = gpu["traj"]; last = tr[tr["step"] == tr["step"].max()]
    np.testing.assert_array_equal(pop["id"], last["id"])
    np.testing.assert_array_equal(pop["y"], last["cols"][:, :nv])
    np.testing.assert_array_equal(pop["aux"], last["cols"][:, nv:nv + na])
    np.testing.assert_array_equal(pop["psi"], last["cols"][:, nv + na:nv + na + npsi])
    if ne:
        np.testing.assert_array_equal(pop["expr"], last["cols"][:, nv + na + npsi + nm:])
    assert max_id >= (pop["id"].max() + 1 if len(pop) else 0)


@pytest.mark.parametrize("fixture,opts,ts,rtol", [
    ("kuramoto", {"meth": "dde", "h0": 0.05, "res": 9, "ps": 300}, (0, 2.0), 1e-10),
    ("kuramoto", {"meth": "fe", "h0": 0.02, "res": 5, "ps": 65}, (0, 1.0), 1e-10),
    ("kuramoto", {"meth": "rkck", "res": 6, "ps": 100}, (0, 5.0), 1e-9),
    ("kuramoto", {"meth": "rk8pd", "res": 4, "ps": 33, "abstol": 1e-9, "reltol": 1e-9}, (0, 3.0), 1e-9),
    ("nbody", {"meth": "dde", "h0": 0.02, "res": 3, "ps": 50, "hmin": 0.0}, (0, 0.08), 1e-8),
    ("nbody", {"meth": "rkf45", "h0": 0.01, "res": 3, "ps": 300, "hmin": 0.0, "abstol": 1e-8, "reltol": 1e-8}, (0, 0.05), 1e-8),
])
def test_patched_reference_binary_gpu_equals_cpu(model_lib, fixture, opts, ts, rtol):
    cpu, gpu, dims = _both(model_lib, fixture, opts, ts)
    _check_files(cpu, gpu, dims, rtol)
    _check_mirror(gpu, dims)


def test_patched_reference_binary_ssa_events(model_lib):
    """birth / death through the patched binary (SSA).  The host-loop draws (SSA_timestep, the event pick) and the first
    population come from the replayed serial stream, so the run agrees with the CPU path up to and including the time and kind of
    the first event; the initial conditions of particles born on the device are keyed Philox draws (the reference draws them from
    libc rand()), so from there on the comparison is with the library's own driver called directly (odx_odesolver through ctypes,
    itself checked against the restatement in tests/test_gpu_events.py): byte-identical files.  SIM->pop / SIM->max_id mirror the
    device population, ids included."""
    from odexp_b200 import capi
    opts = {"meth": "fe", "h0": 0.01, "res": 31, "ps": 20, "ssahmin": 0.0}
    cpu, gpu, dims = _both(model_lib, "lineage", opts, (0, 6))
    a, b = cpu["stats"], gpu["stats"]
    assert len(np.unique(b["n"])) > 1
    k = int(np.argmax(b["event"][20:, 1] != 0)) + 20                    # first event row after the 20 initial rows
    for f in ("n", "stop_flag", "event"):
        np.testing.assert_array_equal(a[f][:k + 1], b[f][:k + 1])
    np.testing.assert_allclose(b["t"][:k + 1], a["t"][:k + 1], rtol=1e-14)
    np.testing.assert_allclose(b["mfd"][:k + 1], a["mfd"][:k + 1], rtol=0, atol=1e-12)
    pop = capi.Population(model_lib("lineage"), capacity=4 * 20 + 4096)
    direct = capi.odesolver(pop, (0, 6), meth="fe", h0=0.01, res=31, popsize=20, ssahmin=0.0, replay_ic=True)
    pop.close()
    assert direct["stats"].tobytes() == b.tobytes()
    assert direct["traj"].tobytes() == gpu["traj"].tobytes()
    _check_mirror(gpu, dims)


def test_patched_reference_binary_fails_loudly_without_library(model_lib, tmp_path):
    """ODEXP_CUDA_MODEL naming a library that does not exist: an error, not a silent CPU run"""
    from oracle import refrun, port
    if not refrun.have_dropin("kuramoto"):
        pytest.skip("model_gpu.out not built")
    with pytest.raises(Exception):
        refrun.run("kuramoto", {"meth": "dde", "h0": 0.1, "res": 3, "ps": 8}, (0, 0.2), dims=port.DIMS["kuramoto"], exe="model_gpu.out",
                   extra_env={"ODEXP_CUDA_MODEL": str(tmp_path / "libodx_missing.so")})


def test_opt_init_1_continues_from_previous_population(model_lib):
    """`opt init 1` (src/methods_odexp.c:289-316): the next run starts from the particles the previous one left (y, expr, ids) and
    writes event {0, 0, i} on its first rows.  With a fixed-step method two runs back to back equal one run over both intervals."""
    from odexp_b200 import capi
    lib = model_lib("kuramoto")
    n = 777
    whole = capi.Population(lib, capacity=1024)
    ref = capi.odesolver(whole, (0, 2.0), meth="dde", h0=0.05, res=41, popsize=n, replay_ic=True)
    whole.close()
    pop = capi.Population(lib, capacity=1024)
    first = capi.odesolver(pop, (0, 1.0), meth="dde", h0=0.05, res=21, popsize=n, replay_ic=True)
    st = pop.download()
    pop2 = capi.Population(lib, capacity=1024)                       # what the branch does: upload SIM->pop into a fresh handle
    pop2.upload(st["y"], st["expr"], ids=st["id"], max_id=n)
    second = capi.odesolver(pop2, (1.0, 2.0), meth="dde", h0=0.05, res=21, popsize=n, replay_ic=True, init_mode=1)
    pop.close(); pop2.close()
    rows = second["stats"][:n]
    np.testing.assert_array_equal(rows["event"][:, :2], 0)
    np.testing.assert_array_equal(rows["event"][:, 2], np.arange(n))
    assert np.all(rows["stop_flag"] == 1)                             # EVENT_T0
    la = ref["traj"][ref["traj"]["step"] == ref["traj"]["step"].max()]
    lb = second["traj"][second["traj"]["step"] == second["traj"]["step"].max()]
    np.testing.assert_array_equal(la["id"], lb["id"])
    np.testing.assert_allclose(lb["cols"], la["cols"], rtol=1e-12, atol=1e-13)


def test_abort_request_stops_the_run(model_lib):
    """odx_request_abort (the SIGINT path, src/methods_odexp.c:23-30, 451): a long fixed-step run is cut into bounded launches and
    stops early; *t reports the time reached; odx_odesolver leaves the loop like the reference does"""
    import threading, time
    from odexp_b200 import capi
    lib = model_lib("kuramoto")
    n = 1 << 22
    pop = capi.Population(lib, capacity=n)
    pop.init(n, replay=False)
    pop.run_fixed("dde", 0.0, 0.05, 10); pop.sync()
    threading.Timer(0.3, lambda: pop.L.odx_request_abort(pop.h)).start()
    w0 = time.perf_counter()
    t = pop.run_fixed("dde", 0.0, 0.05, 200000)                      # ~17 s of work if it ran to the end
    pop.sync()
    wall = time.perf_counter() - w0
    assert pop.L.odx_abort_requested(pop.h) == 1
    assert 0.0 < t < 0.05 * 200000 and wall < 3.0
    pop.L.odx_clear_abort(pop.h)
    t2 = pop.run_fixed("dde", t, 0.05, 4)
    assert t2 == pytest.approx(t + 0.2)
    pop.close()


def test_reference_signature_entry_points(model_lib, tmp_path):
    """fe_apply / dde_apply / SSA_timestep / SSA_apply_birthdeath / tau_leaping_apply_birthdeath called by the reference's names and
    signatures (include/odx_compat.h; sys->params = the handle) drive the same device population as the odx_* calls: the stepping
    loop of odesolver (src/methods_odexp.c:596-616: `while (t < tnext) status = dde_apply(&sys, &t, tnext, &h, y);`) written against
    them reproduces odx_dde_apply, y[] receives the new state, and the birth / death calls write fwrite_SIM rows."""
    import ctypes as C
    from odexp_b200 import capi

    class GslSystem(C.Structure):
        _fields_ = [("function", C.c_void_p), ("jacobian", C.c_void_p), ("dimension", C.c_size_t), ("params", C.c_void_p)]
    lib = model_lib("kuramoto")
    L = lib.lib
    n = 1000
    a = capi.Population(lib, capacity=1024); a.init(n, replay=True)
    b = capi.Population(lib, capacity=1024); b.init(n, replay=True)
    L.odx_compat_bind.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]; L.odx_compat_bind.restype = None
    L.dde_apply.argtypes = [C.POINTER(GslSystem), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.fe_apply.argtypes = L.dde_apply.argtypes
    L.odx_compat_bind(b.h, 1e-5, 0.1, 1, None)
    sys_ = GslSystem(None, None, n, b.h)
    y = np.zeros(n)
    t, h = C.c_double(0.0), C.c_double(0.1)
    ta, ha = 0.0, 0.1
    while t.value < 0.75:                                               # tnext = 0.75 is not a multiple of h: the last step is clipped
        assert L.dde_apply(C.byref(sys_), C.byref(t), 0.75, C.byref(h), y.ctypes.data_as(C.POINTER(C.c_double))) == 0
        ta, ha = a.dde_apply(ta, 0.75, ha, hmin=1e-5, h0=0.1)
        assert t.value == ta and h.value == ha
    np.testing.assert_array_equal(y, a.download()["y"][:, 0])
    np.testing.assert_array_equal(b.download()["y"][:, 0], y)
    assert L.fe_apply(C.byref(sys_), C.byref(t), 1.0, C.byref(h), None) == 0       # y == NULL: state stays on the device
    a.close(); b.close()

    # birth / death by the reference's names, rows in fwrite_SIM layout
    lib2 = model_lib("vdp_sde")
    L2 = lib2.lib
    p = capi.Population(lib2, capacity=4096); p.init(500, replay=False, seed=5); p.rhs(0.0)
    libc = C.CDLL("libc.so.6"); libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]; libc.fclose.argtypes = [C.c_void_p]
    f = libc.fopen(str(tmp_path / "stats.dat").encode(), b"w")
    L2.odx_compat_bind.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]; L2.odx_compat_bind.restype = None
    L2.odx_compat_bind(p.h, 1e-5, 0.01, 0, f)
    L2.SSA_timestep.restype = C.c_double; L2.SSA_timestep.argtypes = [C.POINTER(C.c_double)]
    L2.SSA_apply_birthdeath.argtypes = [C.c_double, C.c_void_p]; L2.SSA_apply_birthdeath.restype = None
    L2.tau_leaping_apply_birthdeath.argtypes = [C.c_double, C.c_double, C.c_void_p]; L2.tau_leaping_apply_birthdeath.restype = None
    r = C.c_double()
    dt = L2.SSA_timestep(C.byref(r))
    assert r.value == pytest.approx(p.rates()[0]) and dt > 0
    n0 = p.size()[0]
    L2.SSA_apply_birthdeath(0.0, None)
    n1 = p.size()[0]
    assert abs(n1 - n0) == 1
    p.rhs(0.0)
    L2.tau_leaping_apply_birthdeath(0.0, 0.5, None)
    n2 = p.size()[0]
    libc.fclose(f)
    rows = np.fromfile(str(tmp_path / "stats.dat"), dtype=np.dtype([("t", "<f8"), ("n", "<i4"), ("stop_flag", "<i4"), ("event", "<i4", (3,))]))
    assert len(rows) >= 2 and rows["n"][0] == n1 and rows["n"][-1] == n2 and np.all(rows["event"][:, 1] != 0)
    p.close()
