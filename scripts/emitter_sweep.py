"""Run the emitter's differential tests over more seeds than the committed suite (CPU only; needs /root/reference for the classes that
compare with the reference binary).  One class at a time -- the classes share oracle/_ref/:

    python scripts/emitter_sweep.py expr    100 160     # peepholes on random expressions, fast vs plain emission (tests/test_emitter_fuzz.py)
    python scripts/emitter_sweep.py model   20  44      # random %M / %C models vs the reference binary (seeds > 50: couplings read MA() / OA())
    python scripts/emitter_sweep.py array   100 124     # array expressions / auxiliaries / equations
    python scripts/emitter_sweep.py lineage 100 118     # birth / death with inheritance: event rows, ids, states
    python scripts/emitter_sweep.py sde     100 130     # Euler-Maruyama: order of the random draws

Round-2 record: expr 100-159 and 200-299, model 20-43, 100-115 and 200-299, array 32-41, 100-123 and 200-299 (about a quarter skipped: the
reference binary aborts in its loader on long array lines), lineage 100-117 and 200-244 (245: the population of the random model explodes,
the reference binary itself does not finish), sde 100-259: no failure.  Frames from the first one in which the reference solution holds a
value beyond 1e12 (a random right-hand side that blows up) are not compared."""
import os
import pathlib
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import pytest                                   # noqa: E402
import test_emitter_fuzz as F                   # noqa: E402
import test_emitter_vs_reference as T           # noqa: E402


class Req:
    def __init__(self):
        self.f = []

    def addfinalizer(self, f):
        self.f.append(f)


which, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
failed = 0
for seed in range(lo, hi):
    tmp, rq = pathlib.Path(tempfile.mkdtemp(prefix="sweep_")), Req()
    try:
        if which == "expr":
            F.test_peepholes_preserve_the_value_of_random_expressions(seed)
        elif which == "model":
            T.test_generated_model_code_matches_the_reference_binary(tmp, seed, rq)
        elif which == "array":
            T.test_generated_array_model_matches_the_reference_binary(tmp, seed, rq)
        elif which == "lineage":
            T.test_generated_birth_death_model_matches_the_reference_binary(tmp, seed, 1.0 if seed % 2 else 0.0, rq)
        elif which == "sde":
            T.test_generated_sde_model_matches_the_reference_binary(tmp, seed, rq)
        else:
            sys.exit("unknown class " + which)
        print("ok", seed, flush=True)
    except pytest.skip.Exception as e:
        print("skip", seed, str(e)[:100], flush=True)
    except BaseException as e:                  # noqa: B036 -- keep sweeping, report at the end
        failed += 1
        print("FAIL", seed, type(e).__name__, str(e)[:1500], flush=True)
    finally:
        for f in rq.f:
            f()
        shutil.rmtree(tmp, ignore_errors=True)
sys.exit(1 if failed else 0)
