"""Build a model library: .pop -> model.cu (cudagen) -> libodx_<name>.so (nvcc, sm_100a) in-tree.

Counterpart of the reference's `compile_model` (src/odexp:1699-1705: gcc model.c -lgsl -lodexp).  The shared object
holds the generated model, the engine kernels instantiated for it, the C-ABI of include/odx.h and the C host driver
(csrc/odx_host.c).  Built files live in odexp_b200/_build/ (git-ignored, shipped to the GPU box by gpurun).
"""
import hashlib, os, shutil, subprocess
from . import popfile, cudagen

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
BUILD = os.path.join(HERE, "_build")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-fno-gnu-unique", "-Xlinker", "-Bsymbolic", "-diag-suppress", "177,550,2361",
              "-shared"]


def _digest(text):
    h = hashlib.sha1(text.encode("utf8"))
    for f in sorted(os.listdir(CSRC)) + ["../../include/odx.h", "../../include/odx_compat.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def lib_path(name):
    return os.path.join(BUILD, "libodx_%s.so" % name)


def build_model(pop_path=None, text=None, name=None, force=False, verbose=False, extra_flags=()):
    """Generate + compile; returns the path of the shared library.  Raises if nvcc is missing or fails."""
    if text is None:
        with open(pop_path, encoding="utf8") as f:
            text = f.read()
    if name is None:
        name = os.path.splitext(os.path.basename(pop_path))[0]
    model = popfile.parse(text, name=name, base_dir=os.path.dirname(os.path.abspath(pop_path)) if pop_path else ".")
    # ODX_EMIT_PLAIN=1: the model expressions as written (no reciprocal / pow / select / zero-product rewrites, cudagen.py): the escape hatch
    # when a rewrite is suspected; the library is cached under its own digest
    plain = os.environ.get("ODX_EMIT_PLAIN", "0") == "1"
    cu = cudagen.emit_cuda(model, fast_pow=not plain, fast_div=not plain)
    os.makedirs(BUILD, exist_ok=True)
    so, cu_path, stamp = lib_path(name), os.path.join(BUILD, "model_%s.cu" % name), os.path.join(BUILD, "model_%s.stamp" % name)
    dg = _digest(cu + " ".join(extra_flags))
    if not force and os.path.exists(so) and os.path.exists(stamp) and open(stamp).read().strip() == dg:
        return so
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        # a prebuilt library is only trusted when its stamp matches the current .pop, csrc and flags (checked above): a stale one
        # would have other dimensions / another ABI than the parse the caller holds
        raise RuntimeError("nvcc not found and the prebuilt %s is missing or stale (stamp != digest of the current sources)" % so)
    with open(cu_path, "w") as f:
        f.write(cu)
    host_objs = []
    for src in ("odx_host.c", "odx_compat.c"):      # host side in C: the GPU branch of odesolver, the reference-signature stepper entry points
        obj = os.path.join(BUILD, "%s_%s.o" % (src[:-2], name))
        subprocess.run(["gcc", "-std=gnu99", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-I" + INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj], check=True)
        host_objs.append(obj)
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-I" + CSRC, "-I" + INCLUDE, cu_path] + host_objs + ["-o", so, "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for model %s:\n%s" % (name, r.stdout))
    with open(stamp, "w") as f:
        f.write(dg)
    return so
