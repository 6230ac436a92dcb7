"""ctypes binding of the C-ABI in include/odx.h (one shared library per model) + a thin solver wrapper.

This is the Python mirror of the reference's simulation entry point: `Population.odesolver(...)` takes the same
options the reference reads from GOPTS (res, hmin, h0, abstol, reltol, meth, popsize, ssahmin, aleap, seed) and
writes the same `.odexp/stats.dat` / `.odexp/traj.dat` files (layouts: src/datastruct.c:643-674), through the C host
driver `odx_odesolver` (csrc/odx_host.c).  There is no CPU fallback: loading fails loudly when the CUDA library is
missing, and every call fails when no B200 is present.
"""
import ctypes as C
import os, tempfile
import numpy as np
from . import build as _build
from . import popfile

METHODS = {"fe": 0, "dde": 1, "rkf45": 2, "rkck": 3, "rk4": 4, "iteration": 5, "rk2": 6, "rk8pd": 7}


class OdxError(RuntimeError):
    pass


def method_code(meth):
    """`opt meth` -> ODX_METH_* (include/odx.h); the implicit GSL steppers of src/methods_odexp.c:154-208 are not built (DESIGN.md section 6:
    the drop-in binary keeps them on the reference's CPU path)"""
    try:
        return METHODS[meth]
    except KeyError:
        raise OdxError("meth %s: not a stepping method of the CUDA path (%s); bsimp and the other implicit GSL methods stay on the "
                       "reference's CPU path" % (meth, ", ".join(sorted(METHODS)))) from None


class ModelInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("nvar", C.c_int), ("nexpr", C.c_int), ("naux", C.c_int), ("npsi", C.c_int),
                ("nmfd", C.c_int), ("npar", C.c_int), ("has_birth", C.c_int), ("has_death", C.c_int), ("has_repli", C.c_int),
                ("rhs_uses_rng", C.c_int), ("has_pairwise", C.c_int), ("has_lowrank", C.c_int),
                ("var_names", C.POINTER(C.c_char_p)), ("expr_names", C.POINTER(C.c_char_p)), ("aux_names", C.POINTER(C.c_char_p)),
                ("psi_names", C.POINTER(C.c_char_p)), ("mfd_names", C.POINTER(C.c_char_p)), ("par_names", C.POINTER(C.c_char_p)),
                ("par_defaults", C.POINTER(C.c_double)), ("has_stopif", C.c_int), ("ndelay", C.c_int)]


class EventRow(C.Structure):
    _fields_ = [("n", C.c_int32), ("stop_flag", C.c_int32), ("event", C.c_int32 * 3)]


class SolverOpts(C.Structure):
    _fields_ = [("method", C.c_int), ("res", C.c_int), ("hmin", C.c_double), ("h0", C.c_double), ("abstol", C.c_double),
                ("reltol", C.c_double), ("ssahmin", C.c_double), ("aleap", C.c_double), ("popsize", C.c_int), ("seed", C.c_uint32),
                ("replay_ic", C.c_int), ("traj_stride", C.c_int), ("outdir", C.c_char_p), ("tspan", C.POINTER(C.c_double)),
                ("tspan_len", C.c_int), ("mu", C.POINTER(C.c_double)),
                ("quick", C.c_int), ("plot_x", C.c_int), ("plot_y", C.c_int), ("plot_z", C.c_int), ("particle", C.c_int),
                ("init_mode", C.c_int)]


class SolverReport(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("solver_ms", C.c_double), ("steps", C.c_int64), ("particle_steps", C.c_double),
                ("status", C.c_int), ("final_n", C.c_int64), ("hmin_reached", C.c_int), ("hmin_t", C.c_double)]


# every symbol include/odx.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = ["odx_model", "odx_create", "odx_destroy", "odx_last_error", "odx_sync", "odx_set_params", "odx_set_seed",
               "odx_set_tolerances", "odx_comm_unique_id", "odx_comm_init", "odx_pop_upload", "odx_pop_download", "odx_pop_size",
               "odx_pop_init", "odx_pop_reinit", "odx_stop_flag", "odx_rhs", "odx_meanfields", "odx_rates", "odx_fe_apply", "odx_dde_apply", "odx_evolve_reset",
               "odx_evolve_apply", "odx_evolve_stats", "odx_run_fixed", "odx_ssa_apply", "odx_tauleap_apply", "odx_draw_uniform",
               "odx_draw_poisson", "odx_odesolver", "odx_event_record", "odx_event_elapsed_ms", "odx_launch_count",
               "odx_set_option", "odx_peer_export", "odx_peer_attach", "odx_comm_rank", "odx_iteration_apply",
               "odx_pop_download_strided", "odx_tauleap_apply_view", "odx_pop_find_id", "odx_lowrank_order",
               "odx_request_abort", "odx_abort_requested", "odx_clear_abort", "odx_frame_begin", "odx_frame_end", "odx_set_event_type"]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class ModelLib:
    """A compiled model library.  `ModelLib.from_pop(path)` generates + compiles (or reuses the in-tree build)."""

    def __init__(self, so_path):
        if not os.path.exists(so_path):
            raise OdxError("model library %s not built: the CUDA extension is required (no CPU fallback)" % so_path)
        self.path = so_path
        self.lib = L = C.CDLL(so_path, mode=C.RTLD_LOCAL)
        H = C.c_void_p
        L.odx_model.restype = C.POINTER(ModelInfo)
        L.odx_last_error.restype = C.c_char_p; L.odx_last_error.argtypes = [H]
        L.odx_create.argtypes = [C.POINTER(H), C.c_int, C.c_int64]
        L.odx_destroy.argtypes = [H]; L.odx_destroy.restype = None
        L.odx_sync.argtypes = [H]
        L.odx_set_params.argtypes = [H, C.POINTER(C.c_double), C.c_int]
        L.odx_set_seed.argtypes = [H, C.c_uint64]
        L.odx_set_event_type.argtypes = [H, C.c_int, C.c_double]
        L.odx_set_tolerances.argtypes = [H, C.c_double, C.c_double]
        L.odx_set_option.argtypes = [H, C.c_char_p, C.c_double]
        L.odx_peer_export.argtypes = [H, C.c_void_p]
        L.odx_peer_attach.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
        L.odx_comm_rank.argtypes = [H, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.odx_comm_unique_id.argtypes = [C.c_void_p]
        L.odx_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
        L.odx_pop_upload.argtypes = [H, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int32]
        L.odx_pop_download.argtypes = [H, C.c_int64, C.c_int64] + [C.POINTER(C.c_double)] * 4 + [C.POINTER(C.c_int32)] + [C.POINTER(C.c_double)] * 2
        L.odx_pop_download_strided.argtypes = [H, C.c_int64, C.c_int64, C.c_int64] + [C.POINTER(C.c_double)] * 4 + [C.POINTER(C.c_int32)]
        L.odx_pop_size.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.odx_pop_find_id.argtypes = [H, C.c_int32, C.POINTER(C.c_int64)]
        L.odx_pop_init.argtypes = [H, C.c_int64, C.c_double, C.c_uint32, C.c_int]
        L.odx_pop_reinit.argtypes = [H, C.c_double, C.c_int]
        L.odx_stop_flag.argtypes = [H, C.POINTER(C.c_int)]
        L.odx_rhs.argtypes = [H, C.c_double]
        L.odx_meanfields.argtypes = [H, C.POINTER(C.c_double)]
        L.odx_rates.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.odx_lowrank_order.argtypes = [H, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        for f in (L.odx_fe_apply, L.odx_dde_apply):
            f.argtypes = [H, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int]
        L.odx_iteration_apply.argtypes = [H, C.POINTER(C.c_double), C.c_int]
        L.odx_evolve_reset.argtypes = [H]
        L.odx_evolve_apply.argtypes = [H, C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
        L.odx_evolve_stats.argtypes = [H] + [C.POINTER(C.c_uint64)] * 3
        L.odx_run_fixed.argtypes = [H, C.c_int, C.POINTER(C.c_double), C.c_double, C.c_int64]
        L.odx_ssa_apply.argtypes = [H, C.c_double, C.c_double, C.POINTER(EventRow)]
        L.odx_tauleap_apply.argtypes = [H, C.c_double, C.c_double, C.c_uint64, C.POINTER(EventRow), C.c_int64, C.POINTER(C.c_int64)]
        L.odx_tauleap_apply_view.argtypes = [H, C.c_double, C.c_double, C.c_uint64, C.POINTER(C.POINTER(EventRow)), C.POINTER(C.c_int64)]
        L.odx_draw_uniform.argtypes = [H, C.c_int]; L.odx_draw_uniform.restype = C.c_double
        L.odx_draw_poisson.argtypes = [H, C.c_double, C.c_int]; L.odx_draw_poisson.restype = C.c_long
        L.odx_odesolver.argtypes = [H, C.POINTER(SolverOpts), C.POINTER(SolverReport)]
        L.odx_event_record.argtypes = [H, C.c_int]
        L.odx_event_elapsed_ms.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.odx_launch_count.argtypes = [H]; L.odx_launch_count.restype = C.c_uint64
        L.odx_request_abort.argtypes = [H]; L.odx_request_abort.restype = None
        L.odx_clear_abort.argtypes = [H]; L.odx_clear_abort.restype = None
        L.odx_abort_requested.argtypes = [H]
        mi = L.odx_model().contents
        self.info = mi
        self.name = mi.name.decode()
        self.nvar, self.nexpr, self.naux, self.npsi, self.nmfd, self.npar = mi.nvar, mi.nexpr, mi.naux, mi.npsi, mi.nmfd, mi.npar
        self.par_defaults = np.array([mi.par_defaults[i] for i in range(mi.npar)], dtype=np.float64)

    @classmethod
    def from_pop(cls, pop_path, name=None, force=False):
        lib = cls(_build.build_model(pop_path, name=name, force=force))
        lib.model = popfile.parse_file(pop_path)
        return lib

    def dims(self):
        return dict(nvar=self.nvar, nexpr=self.nexpr, naux=self.naux, npsi=self.npsi, nmfd=self.nmfd)


class Population:
    """Device-resident population of one model on one GPU (handle of the C-ABI)."""

    def __init__(self, lib: ModelLib, capacity: int, device: int = 0):
        self.lib, self.L = lib, lib.lib
        self.h = C.c_void_p()
        st = self.L.odx_create(C.byref(self.h), device, int(capacity))
        if st != 0:
            msg = self.L.odx_last_error(self.h).decode() if self.h else "odx_create failed"
            if self.h:
                self.L.odx_destroy(self.h); self.h = None
            raise OdxError("odx_create: status %d: %s" % (st, msg))

    def close(self):
        if getattr(self, "h", None):
            self.L.odx_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st, what):
        if st != 0:
            raise OdxError("%s: status %d: %s" % (what, st, self.L.odx_last_error(self.h).decode()))

    # -- configuration
    def set_params(self, mu):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        self._ck(self.L.odx_set_params(self.h, _dp(mu), len(mu)), "odx_set_params")

    def set_seed(self, seed):
        self._ck(self.L.odx_set_seed(self.h, int(seed)), "odx_set_seed")

    def set_tolerances(self, abstol, reltol):
        self._ck(self.L.odx_set_tolerances(self.h, abstol, reltol), "odx_set_tolerances")

    def set_option(self, name, value):
        self._ck(self.L.odx_set_option(self.h, name.encode(), float(value)), "odx_set_option")

    def peer_export(self):
        buf = C.create_string_buffer(64)
        self._ck(self.L.odx_peer_export(self.h, buf), "odx_peer_export")
        return bytes(buf.raw)

    def peer_attach(self, rank, nranks, handles):
        buf = C.create_string_buffer(b"".join(handles), 64 * nranks)
        self._ck(self.L.odx_peer_attach(self.h, rank, nranks, buf), "odx_peer_attach")

    def peer_detach(self):
        r, n = C.c_int(), C.c_int()
        self.L.odx_comm_rank(self.h, C.byref(r), C.byref(n))
        self._ck(self.L.odx_peer_attach(self.h, r.value, n.value, None), "odx_peer_attach(detach)")

    def comm_init(self, rank, nranks, uid):
        buf = C.create_string_buffer(bytes(uid), 128)
        self._ck(self.L.odx_comm_init(self.h, rank, nranks, buf), "odx_comm_init")

    def unique_id(self):
        buf = C.create_string_buffer(128)
        st = self.L.odx_comm_unique_id(buf)
        if st != 0:
            raise OdxError("odx_comm_unique_id: status %d" % st)
        return bytes(buf.raw)

    # -- population
    def init(self, n, t=0.0, seed=3141592, replay=True):
        self._ck(self.L.odx_pop_init(self.h, int(n), float(t), int(seed), 1 if replay else 0), "odx_pop_init")

    def upload(self, y, expr=None, ids=None, max_id=None):
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1, self.lib.nvar)
        n = y.shape[0]
        e = np.ascontiguousarray(expr, dtype=np.float64).reshape(n, self.lib.nexpr) if self.lib.nexpr else None
        i = np.ascontiguousarray(ids, dtype=np.int32) if ids is not None else None
        self._ck(self.L.odx_pop_upload(self.h, n, _dp(y), _dp(e), i.ctypes.data_as(C.POINTER(C.c_int32)) if i is not None else None,
                                       int(max_id if max_id is not None else n)), "odx_pop_upload")

    def size(self):
        a, b, m = C.c_int64(), C.c_int64(), C.c_int32()
        self.L.odx_pop_size(self.h, C.byref(a), C.byref(b), C.byref(m))
        return a.value, b.value, m.value

    def find_id(self, pid):
        """getpar(id): local position of the particle with this id, -1 if it is not on this GPU"""
        pos = C.c_int64()
        self._ck(self.L.odx_pop_find_id(self.h, int(pid), C.byref(pos)), "odx_pop_find_id")
        return pos.value

    def download(self, first=0, count=None, out=None):
        """host copies of the population columns (particle-major); out = the dict of a previous call of the same size, to reuse
        its (already touched) buffers"""
        n = self.size()[0]
        count = n - first if count is None else count
        lb = self.lib
        if out is None or len(out["id"]) != count:
            out = dict(y=np.empty((count, lb.nvar)), expr=np.empty((count, lb.nexpr)), aux=np.empty((count, lb.naux)),
                       psi=np.empty((count, lb.npsi)), id=np.empty(count, dtype=np.int32), death=np.empty(count), repli=np.empty(count))
        self._ck(self.L.odx_pop_download(self.h, first, count, _dp(out["y"]), _dp(out["expr"]), _dp(out["aux"]), _dp(out["psi"]),
                                         out["id"].ctypes.data_as(C.POINTER(C.c_int32)), _dp(out["death"]), _dp(out["repli"])),
                 "odx_pop_download")
        return out

    # -- evaluation / stepping
    def rhs(self, t=0.0):
        self._ck(self.L.odx_rhs(self.h, float(t)), "odx_rhs")

    def meanfields(self):
        m = np.zeros(max(1, self.lib.nmfd))
        self._ck(self.L.odx_meanfields(self.h, _dp(m)), "odx_meanfields")
        return m[:self.lib.nmfd]

    def lowrank_order(self):
        """(p, centre, half-range) chosen by the adaptive low-rank expansion in the last RHS evaluation"""
        p, m, r = C.c_int(), C.c_double(), C.c_double()
        self._ck(self.L.odx_lowrank_order(self.h, C.byref(p), C.byref(m), C.byref(r)), "odx_lowrank_order")
        return p.value, m.value, r.value

    def rates(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.odx_rates(self.h, C.byref(a), C.byref(b)), "odx_rates")
        return a.value, b.value

    def dde_apply(self, t, tnext, h, hmin=1e-5, h0=0.1, want_outputs=True):
        tt, hh = C.c_double(t), C.c_double(h)
        self._ck(self.L.odx_dde_apply(self.h, C.byref(tt), tnext, C.byref(hh), hmin, h0, int(want_outputs)), "odx_dde_apply")
        return tt.value, hh.value

    def fe_apply(self, t, tnext, h, hmin=1e-5, h0=0.1, want_outputs=True):
        tt, hh = C.c_double(t), C.c_double(h)
        self._ck(self.L.odx_fe_apply(self.h, C.byref(tt), tnext, C.byref(hh), hmin, h0, int(want_outputs)), "odx_fe_apply")
        return tt.value, hh.value

    def evolve_apply(self, method, t, t1, h):
        tt, hh = C.c_double(t), C.c_double(h)
        self._ck(self.L.odx_evolve_apply(self.h, method_code(method), C.byref(tt), t1, C.byref(hh)), "odx_evolve_apply")
        return tt.value, hh.value

    def evolve_reset(self):
        self._ck(self.L.odx_evolve_reset(self.h), "odx_evolve_reset")

    def evolve_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.L.odx_evolve_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def run_fixed(self, method, t, h, nsteps):
        tt = C.c_double(t)
        self._ck(self.L.odx_run_fixed(self.h, method_code(method), C.byref(tt), h, int(nsteps)), "odx_run_fixed")
        return tt.value

    def event_record(self, slot):
        self._ck(self.L.odx_event_record(self.h, slot), "odx_event_record")

    def event_elapsed_ms(self, a, b):
        ms = C.c_double()
        self._ck(self.L.odx_event_elapsed_ms(self.h, a, b, C.byref(ms)), "odx_event_elapsed_ms")
        return ms.value

    def launch_count(self):
        return int(self.L.odx_launch_count(self.h))

    def upload_ptr(self, n, y_ptr, expr_ptr, max_id=None):
        """upload from raw host pointers (e.g. pinned torch tensors): particle-major y[n*nvar], expr[n*nexpr]"""
        self._ck(self.L.odx_pop_upload(self.h, int(n), C.cast(y_ptr, C.POINTER(C.c_double)),
                                       C.cast(expr_ptr, C.POINTER(C.c_double)) if expr_ptr else None, None,
                                       int(max_id if max_id is not None else n)), "odx_pop_upload")

    def sync(self):
        self._ck(self.L.odx_sync(self.h), "odx_sync")

    def ssa_apply(self, t, u):
        row = EventRow()
        self._ck(self.L.odx_ssa_apply(self.h, t, u, C.byref(row)), "odx_ssa_apply")
        return (row.n, row.stop_flag, tuple(row.event))

    def tauleap_apply(self, t, dt, leap_index, max_rows=None, as_array=False):
        """one tau-leaping sweep; returns the stats.dat event rows (list of (n, stop_flag, (parent, +-1, child)), or with
        as_array=True a numpy structured array: no per-row Python work, for large populations)"""
        nr = C.c_int64()
        if as_array and max_rows is None:
            # rows stay in the library's pinned buffer; the array is a copy of exactly the rows of this sweep
            view = C.POINTER(EventRow)()
            self._ck(self.L.odx_tauleap_apply_view(self.h, t, dt, int(leap_index), C.byref(view), C.byref(nr)), "odx_tauleap_apply_view")
            dt_row = np.dtype([("n", "<i4"), ("stop_flag", "<i4"), ("event", "<i4", (3,))])
            assert dt_row.itemsize == C.sizeof(EventRow)
            if nr.value == 0:
                return np.zeros(0, dtype=dt_row)
            out = np.empty(nr.value, dtype=dt_row)
            C.memmove(out.ctypes.data, view, nr.value * dt_row.itemsize)
            return out
        n = self.size()[0]
        max_rows = max_rows or (3 * n + 4096)
        if as_array:
            buf = np.zeros(max_rows, dtype=np.dtype([("n", "<i4"), ("stop_flag", "<i4"), ("event", "<i4", (3,))]))
            assert buf.itemsize == C.sizeof(EventRow)
            rows = buf.ctypes.data_as(C.POINTER(EventRow))
        else:
            rows = (EventRow * max_rows)()
        self._ck(self.L.odx_tauleap_apply(self.h, t, dt, int(leap_index), rows, max_rows, C.byref(nr)), "odx_tauleap_apply")
        if as_array:
            return buf[:nr.value]
        return [(rows[i].n, rows[i].stop_flag, tuple(rows[i].event)) for i in range(nr.value)]


def read_outputs(outdir, dims, read_traj=True):
    """Decode stats.dat / traj.dat written by odx_odesolver (same record layouts as the reference)."""
    nmfd = dims["nmfd"]
    sdt = np.dtype([("t", "<f8"), ("mfd", "<f8", (nmfd,)), ("n", "<i4"), ("stop_flag", "<i4"), ("event", "<i4", (3,))])
    out = {"stats": np.fromfile(os.path.join(outdir, "stats.dat"), dtype=sdt)}
    tp = os.path.join(outdir, "traj.dat")
    if read_traj and os.path.exists(tp):
        nc = dims["nvar"] + dims["naux"] + dims["npsi"] + nmfd + dims["nexpr"]
        tdt = np.dtype([("step", "<u4"), ("id", "<i4"), ("t", "<f8"), ("cols", "<f8", (nc,))])
        out["traj"] = np.fromfile(tp, dtype=tdt)
    qp = os.path.join(outdir, "current.plot")
    if os.path.exists(qp):
        out["quick"] = np.fromfile(qp, dtype="<f8").reshape(-1, 3)      # fwrite_quick rows (x, y, z)
    return out


def odesolver(pop: Population, tspan, meth="rkck", res=201, hmin=1e-5, h0=0.1, abstol=1e-6, reltol=1e-6, popsize=1, ssahmin=0.0,
              aleap=0.01, seed=3141592, replay_ic=True, traj_stride=1, outdir=None, mu=None, read_traj=True, quick=True,
              plot=(-1, 0, 1), particle=0, init_mode=0):
    """Python mirror of `odesolver` (src/methods_odexp.h:13-17) driving the C host loop; returns outputs + report.
    quick / plot / particle: current.plot with `opt x|y|z` as indices (-1 = T; reference defaults T, [0], [1]) and `opt particle`."""
    own = outdir is None
    if own:
        outdir = tempfile.mkdtemp(prefix="odx_out_")
    os.makedirs(outdir, exist_ok=True)
    ts = np.ascontiguousarray(tspan, dtype=np.float64)
    muarr = np.ascontiguousarray(mu, dtype=np.float64) if mu is not None else None
    o = SolverOpts(method_code(meth), int(res), hmin, h0, abstol, reltol, ssahmin, aleap, int(popsize), int(seed), 1 if replay_ic else 0,
                   int(traj_stride), outdir.encode(), _dp(ts), len(ts), _dp(muarr), 1 if quick else 0, int(plot[0]), int(plot[1]),
                   int(plot[2]), int(particle), int(init_mode))
    rep = SolverReport()
    st = pop.L.odx_odesolver(pop.h, C.byref(o), C.byref(rep))
    if st != 0:
        raise OdxError("odx_odesolver: status %d: %s" % (st, pop.L.odx_last_error(pop.h).decode()))
    res_ = read_outputs(outdir, pop.lib.dims(), read_traj=read_traj)
    res_["report"] = dict(total_ms=rep.total_ms, solver_ms=rep.solver_ms, steps=rep.steps, particle_steps=rep.particle_steps,
                          final_n=rep.final_n, hmin_reached=int(rep.hmin_reached), hmin_t=rep.hmin_t)
    res_["outdir"] = outdir
    if own:
        import shutil
        shutil.rmtree(outdir, ignore_errors=True)
    return res_


def solver_args(model, **overrides):
    """keyword arguments of odesolver() from the OPT / TIMESPAN lines of a parsed model file, the way the reference starts a run from them
    (GOPTS, src/options.c:6-79; `-a key=value` on the command line = `overrides`, with the same abbreviations: ps, meth, r)"""
    o = dict(model.options)
    for k, v in overrides.items():
        o[popfile.OPT_ABBREV.get(k, k)] = str(v)
    if o.get("popmode", "population")[:3] == "sin":
        raise OdxError("popmode single: one trajectory, no population -- stays on the reference's CPU path")
    return dict(tspan=tuple(model.timespan), meth=o["solver"], res=int(float(o["res"])), hmin=float(o["hmin"]), h0=float(o["h0"]),
                abstol=float(o["abstol"]), reltol=float(o["reltol"]), popsize=int(float(o["popsize"])), ssahmin=float(o["ssahmin"]),
                aleap=float(o["aleap"]), seed=int(float(o["seed"])), init_mode=int(float(o.get("init", "0"))),
                particle=int(float(o.get("particle", "0"))))


def run_pop(pop_path, capacity=None, device=0, replay_ic=True, **overrides):
    """`odexp model.pop` for the population path: build (or reuse) the model library, size the population from `opt popsize`
    (with room for births), integrate over the file's TIMESPAN with its OPT values; overrides as in solver_args()"""
    lib = ModelLib.from_pop(pop_path)
    kw = solver_args(lib.model, **overrides)
    births = bool(lib.model.birth or lib.model.repli)
    pop = Population(lib, capacity=capacity or (max(4 * kw["popsize"], 1024) if births else kw["popsize"] + 8), device=device)
    try:
        return odesolver(pop, kw.pop("tspan"), replay_ic=replay_ic, **kw)
    finally:
        pop.close()
