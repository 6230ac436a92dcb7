"""Multi-GPU parity, run by `pytest -m gpu` whenever the box shows at least two GPUs (skipped on a single-GPU box): the sharded
run must reproduce the single-GPU run.  Spawns one process per GPU with torch.distributed.run, like the bench."""
import os, subprocess, sys
import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_population_equals_single_gpu(world):
    """scripts/mgpu_check.py: for every model family (mean field through the in-kernel NVLink exchange in its plain / ring / lean
    variants, all-pairs %C through the allgather, low-rank through the moment allreduce, SDE, {vect} mean fields) each rank's
    shard equals the same particles of a single-GPU run, and whole simulations with SSA / tau-leaping write identical
    stats.dat rows on every rank."""
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29700 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "scripts", "mgpu_check.py")], cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-6000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:]
    assert "MISMATCH" not in r.stdout
