#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — build the UNCHANGED reference CPU path into oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may use
anything under oracle/.  This recipe compiles the reference sources *where they lie* under
/root/reference (nothing is copied into the repo), against the declaration-only headers in
oracle/stubinc/, the link stubs in oracle/stubs.c and the clean-room GSL subset oracle/minigsl.c.
Outputs go ONLY to oracle/_ref/ (git-ignored, travels to the GPU box with gpurun):

    oracle/_ref/libodexp_ref.a          main.c methods_odexp.c datastruct.c loader.c options.c
                                        rand_gen.c utils_odexp.c lrexp.c + stubs + minigsl
    oracle/_ref/<fixture>/model.out     reference `perl src/odexp -d -q` output for the fixture, compiled -O3
    oracle/_ref/<fixture>/.odexp/       equations.pop / parameters.pop the binary re-reads at run time

The reference's own build system (cmake, macOS dylibs, libgsl) is NOT used; see DESIGN.md §oracle.
"""
import os, shutil, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ODX_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB_SRCS = ["main", "methods_odexp", "datastruct", "loader", "options", "rand_gen", "utils_odexp", "lrexp"]
# _FORTIFY_SOURCE off: the reference's loader passes over-stated buffer lengths to snprintf/strncat for array entries
# (`expr U[i=0:R] ...`); harmless on its macOS target, aborted by glibc's fortified variants
CFLAGS = ["-std=gnu99", "-O3", "-D_GNU_SOURCE", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-w", "-I" + os.path.join(HERE, "stubinc"),
          "-I" + os.path.join(REF, "src"), "-include", "errno.h", "-include", "gsl/gsl_errno.h", "-include", "bsd_compat.h"]


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def available():
    return os.path.isfile(os.path.join(REF, "src", "methods_odexp.c")) and shutil.which("perl") and shutil.which("gcc")


def fixtures():
    fx = {}
    d = os.path.join(os.path.dirname(HERE), "models")
    for f in sorted(os.listdir(d)):
        if f.endswith(".pop"):
            fx[f[:-4]] = os.path.join(d, f)
    shipped = os.path.join(REF, "examples", "kuramoto.pop")   # BASELINE config 1: "as shipped"
    if os.path.isfile(shipped):
        fx["kuramoto_shipped"] = shipped
    return fx


def build(force=False, verbose=False):
    if not available():
        return False
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    lib = os.path.join(OUT, "libodexp_ref.a")
    objs = []
    for s in LIB_SRCS:
        src, obj = os.path.join(REF, "src", s + ".c"), os.path.join(OUT, "obj", s + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
            run(["gcc"] + CFLAGS + ["-c", src, "-o", obj])
        objs.append(obj)
    for s in ["stubs", "minigsl", "rand_hook"]:
        src, obj = os.path.join(HERE, s + ".c"), os.path.join(OUT, "obj", s + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
            run(["gcc", "-std=gnu99", "-O3", "-Wall", "-I" + os.path.join(HERE, "stubinc"), "-c", src, "-o", obj])
        objs.append(obj)
    newest_obj = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < newest_obj:
        if os.path.exists(lib):
            os.remove(lib)
        run(["ar", "rcs", lib] + objs)
    gobjs = dropin_objects(force)
    for name, pop in fixtures().items():
        d = os.path.join(OUT, name)
        exe = os.path.join(d, "model.out")
        # incremental: a fixture is regenerated when its model file (or a data file next to it), the reference's generator, an object
        # of the library or this recipe is newer than its binaries
        deps = [pop, os.path.join(REF, "src", "odexp"), os.path.abspath(__file__)] + \
               [os.path.join(os.path.dirname(pop), x) for x in os.listdir(os.path.dirname(pop)) if x.startswith(name + ".") and not x.endswith(".pop")]
        stamp = max([newest_obj] + [os.path.getmtime(f) for f in deps])
        outs = [exe] + ([os.path.join(d, "model_gpu.out")] if name in DROPIN_FIXTURES else [])
        gstamp = max([stamp] + [os.path.getmtime(o) for o in gobjs]) if name in DROPIN_FIXTURES else stamp
        if not force and all(os.path.exists(o) for o in outs) and os.path.getmtime(exe) > stamp and os.path.getmtime(outs[-1]) > gstamp \
                and os.path.isdir(os.path.join(d, ".odexp")):
            continue
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        local = os.path.join(d, os.path.basename(pop))
        shutil.copy(pop, local)
        for extra in os.listdir(os.path.dirname(pop)):           # data files a fixture reads at generation time (FILE keyword)
            if extra.startswith(name + ".") and not extra.endswith(".pop"):
                shutil.copy(os.path.join(os.path.dirname(pop), extra), os.path.join(d, extra))
        # reference generator: -d generate only, -q quiet (appends `OPT loudness quiet`)
        run(["perl", os.path.join(REF, "src", "odexp"), "-d", "-q", os.path.basename(pop)], cwd=d)
        # link order: model + whole archive objects first so rand()/srand() from rand_hook.o win over libc
        run(["gcc"] + CFLAGS + ["-I" + os.path.join(d, ".odexp"), os.path.join(d, ".odexp", "model.c")]
            + objs + ["-lm", "-o", exe])
        os.remove(local)
        if verbose:
            print("built", exe)
        # the drop-in test artefact: the same model.c linked against the reference with the B200 branch added to a COPY of
        # methods_odexp.c (integration/apply_gpu_branch.py) + integration/odexp_gpu_branch.c -> model_gpu.out.  It runs the GPU
        # branch when ODEXP_CUDA_MODEL names the model library, and the unchanged CPU path otherwise.
        if name in DROPIN_FIXTURES:
            run(["gcc"] + CFLAGS + ["-I" + os.path.join(d, ".odexp"), os.path.join(d, ".odexp", "model.c")]
                + [o for o in objs if not o.endswith(os.sep + "methods_odexp.o")] + gobjs + ["-lm", "-ldl", "-o", os.path.join(d, "model_gpu.out")])
    return True


def build_one(name, pop_text):
    """one more model.out under oracle/_ref/<name>/ from a .pop text (tests that generate models: tests/test_emitter_vs_reference.py);
    the reference library must have been built (build())"""
    objs = [os.path.join(OUT, "obj", s + ".o") for s in LIB_SRCS + ["stubs", "minigsl", "rand_hook"]]
    if not available() or not all(os.path.exists(o) for o in objs):
        return None
    d = os.path.join(OUT, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    local = os.path.join(d, name + ".pop")
    open(local, "w").write(pop_text)
    run(["perl", os.path.join(REF, "src", "odexp"), "-d", "-q", os.path.basename(local)], cwd=d)
    exe = os.path.join(d, "model.out")
    run(["gcc"] + CFLAGS + ["-I" + os.path.join(d, ".odexp"), os.path.join(d, ".odexp", "model.c")] + objs + ["-lm", "-o", exe])
    os.remove(local)
    return exe


DROPIN_FIXTURES = ("kuramoto", "lineage", "nbody", "vdp_sde")
ROOT = os.path.dirname(HERE)


def dropin_objects(force=False):
    """methods_odexp.c with the GPU branch inserted (a copy under oracle/_ref/obj) + the branch itself (product source)"""
    gen = os.path.join(OUT, "obj", "methods_odexp_gpu.c")
    script = os.path.join(ROOT, "integration", "apply_gpu_branch.py")
    branch = os.path.join(ROOT, "integration", "odexp_gpu_branch.c")
    src = os.path.join(REF, "src", "methods_odexp.c")
    o1, o2 = os.path.join(OUT, "obj", "methods_odexp_gpu.o"), os.path.join(OUT, "obj", "odexp_gpu_branch.o")
    newest = max(os.path.getmtime(f) for f in (script, branch, src, os.path.join(ROOT, "include", "odx.h")))
    if force or not (os.path.exists(o1) and os.path.exists(o2)) or min(os.path.getmtime(o1), os.path.getmtime(o2)) < newest:
        run([sys.executable, script, src, gen, "--diff", os.path.join(OUT, "odesolver_gpu.patch")])
        run(["gcc"] + CFLAGS + ["-c", gen, "-o", o1])
        run(["gcc"] + CFLAGS + ["-I" + os.path.join(ROOT, "include"), "-c", branch, "-o", o2])
    return [o1, o2]


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref:", "built" if ok else "reference or toolchain not present; skipped")
