#!/usr/bin/env python3
"""torchrun helper: Kuramoto N=2^24 strong-scaling step time for a named model-library variant (development tool)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from odexp_b200 import capi, build, dist as odist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 24
for tag in sys.argv[1:]:
    lib = capi.ModelLib(build.lib_path(tag))
    lo, hi = odist.shard_range(n, rank, world)
    pop = capi.Population(lib, capacity=hi - lo, device=local)
    odist.connect(pop)
    pop.init(hi - lo, replay=False)
    t = pop.run_fixed("dde", 0.0, 0.05, 10); pop.sync(); dist.barrier()
    best = 1e9
    for rep in range(3):
        pop.event_record(0); t = pop.run_fixed("dde", t, 0.05, 200); pop.event_record(1)
        ms = torch.tensor([pop.event_elapsed_ms(0, 1) / 200], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = min(best, float(ms.item()))
    if rank == 0:
        print("%-20s world=%d  %.4f ms/step  %.3e particle-steps/s" % (tag, world, best, n / (best * 1e-3)), flush=True)
    pop.close(); dist.barrier()
dist.destroy_process_group()
