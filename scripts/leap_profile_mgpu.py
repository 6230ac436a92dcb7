#!/usr/bin/env python3
"""sharded tau-leaping sweep timing under torchrun (development tool): N = 2^22 in total, device time between events per sweep"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from odexp_b200 import capi, dist as odist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "vdp_sde.pop"))
n = 1 << 22
lo, hi = odist.shard_range(n, rank, world)
pop = capi.Population(lib, capacity=2 * (hi - lo) + 8192, device=local)
odist.connect(pop)
pop.init(hi - lo, replay=False, seed=8)
t = pop.run_fixed("fe", 0.0, 0.001, 20); pop.sync()
for k in range(10):
    pop.rhs(t); pop.sync(); dist.barrier()
    pop.event_record(0)
    rows = pop.tauleap_apply(t, 1.0, k, as_array=True)
    pop.event_record(1)
    ms = pop.event_elapsed_ms(0, 1)
    if rank == 0:
        print("world=%d sweep %d: %d rows, N -> %d, %.3f ms" % (world, k, len(rows), pop.size()[1], ms), flush=True)
pop.close()
dist.destroy_process_group()
