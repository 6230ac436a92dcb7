"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): sharding arithmetic and handle distribution."""
import os, sys
import torch.multiprocessing as mp
from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from odexp_b200 import dist as odist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = odist.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
    handles = odist.all_gather_bytes(bytes([rank]) * 64)
    lo, hi = odist.shard_range(1000003, rank, world)

    class FakePop:                      # records what connect() hands to the C-ABI
        def unique_id(self): return b"U" * 128
        def comm_init(self, r, w, u): self.comm = (r, w, u)
        def peer_export(self): return bytes([rank + 65]) * 64
        def peer_attach(self, r, w, h): self.peers = (r, w, h)
        def peer_detach(self): self.peers = None
    fp = FakePop()
    odist.connect(fp)
    # a loop count derived from rank-local timings (bench.py's untimed continuation): every rank must end up with rank 0's
    steps = odist.agreed(2094 - rank, src=0)
    q.put((rank, uid, handles, lo, hi, fp.comm, fp.peers, steps))
    dist.destroy_process_group()


def test_two_rank_plumbing():
    world, port = 2, 29500 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[7] for r in res] == [2094, 2094]
    res = [r[:7] for r in res]
    for rank, uid, handles, lo, hi, comm, peers in res:
        assert uid == bytes(range(128))
        assert handles == [bytes([0]) * 64, bytes([1]) * 64]
        assert comm == (rank, 2, b"U" * 128)
        assert peers == (rank, 2, [b"A" * 64, b"B" * 64])
    assert res[0][3] == 0 and res[0][4] == res[1][3] and res[1][4] == 1000003      # contiguous cover of [0, N)
    assert (res[0][4] - res[0][3]) - (res[1][4] - res[1][3]) in (0, 1)


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from odexp_b200.dist import shard_range
    for n in (0, 1, 7, 8, 1 << 24, 12345):
        for w in (1, 2, 3, 8):
            prev = 0
            for r in range(w):
                lo, hi = shard_range(n, r, w)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
