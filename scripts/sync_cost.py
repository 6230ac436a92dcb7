#!/usr/bin/env python3
"""Per-stage cost of the in-kernel mean-field exchange (development tool, run under torchrun or alone): Kuramoto RK4 on a
population so small that every CTA has one tile per stage, so a stage is almost only the barrier / exchange."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from odexp_b200 import capi, dist as odist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "kuramoto.pop"))
for n_local in (148 * 1024, 148 * 1536 * 4, 1 << 21):
    pop = capi.Population(lib, capacity=n_local, device=local)
    if world > 1:
        odist.connect(pop)
    pop.init(n_local, replay=False)
    t = pop.run_fixed("dde", 0.0, 0.05, 50); pop.sync()
    if world > 1:
        dist.barrier()
    steps = 2000
    t = pop.run_fixed("dde", t, 0.05, 1)
    pop.event_record(0); t = pop.run_fixed("dde", t, 0.05, steps); pop.event_record(1)
    ms = pop.event_elapsed_ms(0, 1)
    if rank == 0:
        print("world=%d n_local=%d: %.3f us per step = %.3f us per stage" % (world, n_local, 1e3 * ms / steps, 1e3 * ms / steps / 4), flush=True)
    pop.sync()
    if world > 1:
        dist.barrier()
    pop.close()
if world > 1:
    dist.destroy_process_group()
