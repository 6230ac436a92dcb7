"""Do the fixtures still generate the code their built libraries were compiled from?  Compares the digest of what the emitter prints
today (plus csrc/ and the flags, odexp_b200/build.py) with the stamp next to every libodx_<model>.so, without compiling anything.
Used before an emitter change is allowed near the GPU-validated libraries: exit code 1 lists the fixtures that would rebuild."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odexp_b200 import popfile, cudagen, build      # noqa: E402
import __graft_entry__ as g                          # noqa: E402

stale = []
for name in g.MODELS:
    cu = cudagen.emit_cuda(popfile.parse_file(os.path.join(ROOT, "models", name + ".pop")))
    stamp = os.path.join(build.BUILD, "model_%s.stamp" % name)
    if not os.path.exists(stamp) or open(stamp).read().strip() != build._digest(cu):
        stale.append(name)
# the model-language test models that run on the GPU (tests/test_gpu_syntax.py)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_emitter_vs_reference import SYNTAX_MODELS   # noqa: E402
from test_gpu_syntax import CASES                      # noqa: E402
extra = sorted({c[0] for c in CASES})
for k in extra:
    cu = cudagen.emit_cuda(popfile.parse(SYNTAX_MODELS[k][0], name="syntax_" + k))
    stamp = os.path.join(build.BUILD, "model_syntax_%s.stamp" % k)
    if not os.path.exists(stamp) or open(stamp).read().strip() != build._digest(cu):
        stale.append("syntax_" + k)
print("unchanged: %d of %d" % (len(g.MODELS) + len(extra) - len(stale), len(g.MODELS) + len(extra)) + ("; would rebuild: " + " ".join(stale) if stale else ""))
sys.exit(1 if stale else 0)
