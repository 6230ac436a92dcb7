"""TEST INFRASTRUCTURE — ctypes binding of the CPU restatement (oracle/odx_oracle.c, built as oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
`run(...)` mirrors oracle/refrun.run(...) (same option names, same decoded outputs) so a test can compare the
unchanged reference binary, this restatement and the CUDA library with the same code.
"""
import ctypes as C
import os, shutil, subprocess, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
METHODS = {"fe": 0, "dde": 1, "rkf45": 2, "rkck": 3, "rk4": 4, "iteration": 5, "rk2": 6, "rk8pd": 7}
DIMS = {"kuramoto": dict(nvar=1, naux=2, npsi=0, nmfd=3, nexpr=1), "nbody": dict(nvar=6, naux=0, npsi=3, nmfd=0, nexpr=0),
        "vdp_sde": dict(nvar=2, naux=0, npsi=0, nmfd=0, nexpr=0),
        "kuramoto_lowrank": dict(nvar=1, naux=2, npsi=1, nmfd=2, nexpr=17),
        "kuramoto_lowrank64": dict(nvar=1, naux=2, npsi=1, nmfd=2, nexpr=129),
        "logistic_map": dict(nvar=1, naux=1, npsi=0, nmfd=1, nexpr=1),
        "reset_decay": dict(nvar=1, naux=1, npsi=0, nmfd=1, nexpr=1),
        "ring_diffusion": dict(nvar=2, naux=1, npsi=0, nmfd=1, nexpr=1),
        "vect_meanfield": dict(nvar=1, naux=1, npsi=0, nmfd=3, nexpr=0),
        "chain_array": dict(nvar=6, naux=6, npsi=0, nmfd=2, nexpr=4),
        "fun_block": dict(nvar=3, naux=5, npsi=0, nmfd=1, nexpr=6),
        "helper_calls": dict(nvar=4, naux=13, npsi=0, nmfd=1, nexpr=6),
        "delay_pop": dict(nvar=1, naux=1, npsi=0, nmfd=1, nexpr=1),
        "kuramoto_lrexp": dict(nvar=1, naux=1, npsi=1, nmfd=2, nexpr=1),
        "pair_aux": dict(nvar=2, naux=3, npsi=2, nmfd=1, nexpr=1),
        "repel2d": dict(nvar=4, naux=0, npsi=2, nmfd=2, nexpr=0),
        "lineage": dict(nvar=1, naux=1, npsi=0, nmfd=2, nexpr=4),
        "table_drive": dict(nvar=1, naux=1, npsi=0, nmfd=1, nexpr=0),
        "interp_drive": dict(nvar=1, naux=2, npsi=0, nmfd=1, nexpr=1),
        "lag_chain": dict(nvar=6, naux=1, npsi=0, nmfd=1, nexpr=0),
        "nbody_const": dict(nvar=6, naux=0, npsi=3, nmfd=0, nexpr=0),
        "kuramoto_lowrank_aux": dict(nvar=1, naux=2, npsi=1, nmfd=2, nexpr=7)}


class Opts(C.Structure):
    _fields_ = [("method", C.c_int), ("res", C.c_int), ("hmin", C.c_double), ("h0", C.c_double), ("abstol", C.c_double),
                ("reltol", C.c_double), ("ssahmin", C.c_double), ("aleap", C.c_double), ("popsize", C.c_int), ("seed", C.c_uint32),
                ("rng_mode", C.c_int), ("write_traj", C.c_int), ("outdir", C.c_char_p), ("tspan", C.POINTER(C.c_double)),
                ("tspan_len", C.c_int), ("mu", C.POINTER(C.c_double)),
                ("plot_x", C.c_int), ("plot_y", C.c_int), ("plot_z", C.c_int), ("particle", C.c_int)]


class Report(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("solver_ms", C.c_double), ("function_ms", C.c_double), ("steps", C.c_int64),
                ("particle_steps", C.c_double), ("final_n", C.c_int64), ("status", C.c_int), ("rand_draws", C.c_uint64),
                ("rhs_evals", C.c_uint64)]


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("odx_oracle.c", "models.c", "minigsl.c", "lowrank.c", "odx_oracle.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    subprocess.run(["gcc", "-std=gnu99", "-O3", "-fPIC", "-Wall", "-Wno-unused-parameter", "-I" + os.path.join(HERE, "stubinc"),
                    "-I" + HERE, "-shared"] + srcs[:4] + ["-lm", "-o", LIB], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_odesolver.argtypes = [C.c_char_p, C.POINTER(Opts), C.POINTER(Report)]
    return _lib


def run(model, opts=None, timespan=(0.0, 1.0), rng_mode=0, read_traj=True, mu=None, outdir=None, plot=(-1, 0, 1), particle=0):
    """plot = (x, y, z) axis indices of current.plot: -1 = T, then dynamical variables, auxiliaries, couplings, mean fields"""
    o = dict(meth="rkck", res=201, hmin=1e-5, h0=0.1, abstol=1e-6, reltol=1e-6, ssahmin=0.0, aleap=0.01, ps=1, seed=3141592)
    o.update(opts or {})
    own = outdir is None
    if own:
        outdir = tempfile.mkdtemp(prefix="odxorc_")
    ts = np.ascontiguousarray(timespan, dtype=np.float64)
    muarr = np.ascontiguousarray(mu, dtype=np.float64) if mu is not None else None
    op = Opts(METHODS[o["meth"]], int(o["res"]), float(o["hmin"]), float(o["h0"]), float(o["abstol"]), float(o["reltol"]),
              float(o["ssahmin"]), float(o["aleap"]), int(o["ps"]), int(o["seed"]), int(rng_mode), 1 if read_traj else 0,
              outdir.encode(), ts.ctypes.data_as(C.POINTER(C.c_double)), len(ts),
              muarr.ctypes.data_as(C.POINTER(C.c_double)) if muarr is not None else None,
              int(plot[0]), int(plot[1]), int(plot[2]), int(particle))
    rep = Report()
    st = lib().orc_odesolver(model.encode(), C.byref(op), C.byref(rep))
    if st != 0:
        raise RuntimeError("orc_odesolver status %d" % st)
    d = DIMS[model]
    sdt = np.dtype([("t", "<f8"), ("mfd", "<f8", (d["nmfd"],)), ("n", "<i4"), ("stop_flag", "<i4"), ("event", "<i4", (3,))])
    out = {"stats": np.fromfile(os.path.join(outdir, "stats.dat"), dtype=sdt)}
    if read_traj:
        nc = d["nvar"] + d["naux"] + d["npsi"] + d["nmfd"] + d["nexpr"]
        tdt = np.dtype([("step", "<u4"), ("id", "<i4"), ("t", "<f8"), ("cols", "<f8", (nc,))])
        out["traj"] = np.fromfile(os.path.join(outdir, "traj.dat"), dtype=tdt)
    out["quick"] = np.fromfile(os.path.join(outdir, "current.plot"), dtype="<f8").reshape(-1, 3)
    out["report"] = {k: getattr(rep, k) for k, _ in Report._fields_}
    if own:
        shutil.rmtree(outdir, ignore_errors=True)
    return out


def time_kuramoto_rk4(n, nsteps, h):
    r = run("kuramoto", {"meth": "dde", "h0": h, "res": 2, "ps": n, "hmin": 1e-9}, timespan=(0.0, h * nsteps), read_traj=False)
    return n * nsteps / (r["report"]["solver_ms"] * 1e-3)
