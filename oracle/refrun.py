"""TEST INFRASTRUCTURE — run a prebuilt reference binary (oracle/_ref/<fixture>/model.out) headless.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
The binary is the UNCHANGED reference (see oracle/build_ref.py); it re-reads `.odexp/parameters.pop` at start-up
(src/main.c:163-337), so options and the time span are overridden by appending `OPT k v` lines (later lines
win, src/loader.c:239-275) and rewriting the TIMESPAN line.  Output record layouts: src/datastruct.c:643-674.
"""
import os, re, shutil, subprocess, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def have(fixture):
    return os.path.isfile(os.path.join(REFDIR, fixture, "model.out"))


def _count(text, kw):
    return len(re.findall(r"^\s*%s" % kw, text, flags=re.I | re.M))


def stats_dtype(nmfd):
    return np.dtype([("t", "<f8"), ("mfd", "<f8", (nmfd,)), ("n", "<i4"), ("stop_flag", "<i4"), ("event", "<i4", (3,))])


def traj_dtype(ncol_doubles):
    # unsigned step; int id; double t; y[nvar] aux[naux] psi[npsi] mfd[nmfd] expr[nexpr]
    return np.dtype([("step", "<u4"), ("id", "<i4"), ("t", "<f8"), ("cols", "<f8", (ncol_doubles,))])


def have_dropin(fixture):
    return os.path.isfile(os.path.join(REFDIR, fixture, "model_gpu.out"))


def run(fixture, opts=None, timespan=None, dims=None, keep=False, rand_log=False, timeout=3600, read_traj=True, exe="model.out", extra_env=None):
    """Run fixture with option overrides.  dims = dict(nvar,naux,npsi,nmfd,nexpr) to decode traj.dat.
    Returns dict(stats=..., traj=..., solver_ms=..., total_ms=..., function_ms=..., stdout=..., rand=...)."""
    src = os.path.join(REFDIR, fixture)
    if not have(fixture):
        raise FileNotFoundError("oracle/_ref/%s not built (python oracle/build_ref.py)" % fixture)
    work = tempfile.mkdtemp(prefix="odxref_")
    try:
        os.makedirs(os.path.join(work, ".odexp"))
        for f in ("equations.pop", "parameters.pop"):
            shutil.copy(os.path.join(src, ".odexp", f), os.path.join(work, ".odexp", f))
        ppath = os.path.join(work, ".odexp", "parameters.pop")
        text = open(ppath).read()
        if timespan is not None:
            ts = "timespan " + " ".join(repr(float(x)) for x in timespan)
            text, n = re.subn(r"^\s*times\w*\s+.*$", ts, text, flags=re.I | re.M)
            if n == 0:
                text += "\n" + ts + "\n"
        # every solver option is written out (the restatement's defaults = src/options.c's for the ones not given), so that a
        # fixture's own `opt` lines cannot make the two sides of a comparison run with different values
        full = dict(meth="rkck", res=201, hmin=1e-5, h0=0.1, abstol=1e-6, reltol=1e-6, ssahmin=0.0, aleap=0.01, seed=3141592)
        full.update(opts or {})
        for k, v in (full.items() if opts else ()):          # opts = {} / None: the fixture exactly as written
            text += "\nOPT %s %s\n" % (k, v)
        text += "\nOPT pr 0\n"
        open(ppath, "w").write(text)
        env = dict(os.environ)
        if rand_log:
            env["ODX_RAND_LOG"] = os.path.join(work, "rand.log")
        env.update(extra_env or {})
        # unbuffered stdout: the timing line would otherwise be lost in the reference's exit-time SIGSEGV
        cmd = ([shutil.which("stdbuf"), "-o0"] if shutil.which("stdbuf") else []) + [os.path.join(src, exe)]
        r = subprocess.run(cmd, cwd=work, env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, timeout=timeout)
        # quiet mode ends in the reference's teardown pclose(NULL) SIGSEGV (src/main.c:1739) AFTER outputs are closed.
        out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout.decode("utf8", "replace"))
        res = {"stdout": out, "returncode": r.returncode}
        m = re.search(r"total:\s*([0-9.eE+-]+) msec,\s*solver:\s*([0-9.eE+-]+) msec,\s*function:\s*([0-9.eE+-]+) msec", out)
        if m:
            res["total_ms"], res["solver_ms"], res["function_ms"] = (float(m.group(i)) for i in (1, 2, 3))
        if dims is None:
            eq = open(os.path.join(work, ".odexp", "equations.pop")).read()
            dims = {"nmfd": _count(eq, "%M")}
        nmfd = dims["nmfd"]
        res["stats"] = np.fromfile(os.path.join(work, ".odexp", "stats.dat"), dtype=stats_dtype(nmfd))
        if read_traj and all(k in dims for k in ("nvar", "naux", "npsi", "nexpr")):
            nc = dims["nvar"] + dims["naux"] + dims["npsi"] + nmfd + dims["nexpr"]
            res["traj"] = np.fromfile(os.path.join(work, ".odexp", "traj.dat"), dtype=traj_dtype(nc))
        qp = os.path.join(work, ".odexp", "current.plot")
        if os.path.exists(qp):
            res["quick"] = np.fromfile(qp, dtype="<f8").reshape(-1, 3)
        if rand_log and os.path.exists(env["ODX_RAND_LOG"]):
            res["rand"] = np.fromfile(env["ODX_RAND_LOG"], dtype="<i4")
        dp = (extra_env or {}).get("ODEXP_CUDA_DUMP_POP")
        if dp and os.path.exists(os.path.join(work, dp) if not os.path.isabs(dp) else dp):
            res["pop_dump"] = open(os.path.join(work, dp) if not os.path.isabs(dp) else dp, "rb").read()
        if keep:
            res["workdir"] = work
        return res
    finally:
        if not keep:
            shutil.rmtree(work, ignore_errors=True)
