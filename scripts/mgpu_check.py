#!/usr/bin/env python3
"""multi-GPU correctness check (run under torchrun, one rank per GPU): the sharded run must reproduce the single-GPU run.
  torchrun --nproc-per-node 2 scripts/mgpu_check.py
Each rank integrates its shard; rank 0 also integrates the whole population alone and compares its shard + mean fields."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from odexp_b200 import capi, dist as odist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
CASES = [("kuramoto", 100003, "dde", 20, 1e-12, {}), ("kuramoto", 100003, "dde", 20, 1e-12, {"tma": 1, "lean": 0}),
         ("kuramoto", 100003, "dde", 20, 1e-12, {"tma": 1, "lean": 1}),          # ring / lean step kernels with the NVLink exchange
         ("nbody", 3001, "dde", 2, 1e-12, {}), ("nbody", 2048, "rkf45", 3, 1e-9, {}),
         # symmetric all-pairs kernel on a sharded population (half of every cross-rank tile pair per rank, sums sent to the owner):
         # forced at a ragged size against the ordered single-GPU kernel, automatic at 40 tiles against the symmetric single-GPU one
         ("nbody", 5003, "dde", 2, 1e-12, {"pair_sym": 1}), ("nbody", 40001, "rkf45", 2, 1e-9, {}),
         ("vdp_sde", 50001, "fe", 10, 1e-12, {}), ("vdp_sde", 50001, "fe", 10, 1e-12, {"tma": 1}), ("kuramoto_lowrank", 4097, "dde", 3, 1e-10, {}),
         ("kuramoto_lowrank64", 4099, "dde", 2, 1e-10, {}),
         # model-language coverage added late in round 1: MA()/OA() in %C, low-rank in an auxiliary variable, unweighted lrexp,
         # array variables + MA() in %M, {vect} mean fields
         ("pair_aux", 3001, "dde", 3, 1e-12, {}), ("kuramoto_lowrank_aux", 4097, "dde", 3, 1e-10, {}), ("kuramoto_lrexp", 4097, "dde", 3, 1e-10, {}),
         ("chain_array", 50001, "dde", 5, 1e-12, {}), ("vect_meanfield", 50001, "fe", 5, 1e-12, {})]
ONLY = os.environ.get("MGPU_ONLY", "")               # "leap": only the tau-leaping sections (development runs)
QUICK = bool(os.environ.get("MGPU_QUICK"))          # a subset for boxes where GPU time is dear (8 ranks): mean field, symmetric %C, one sweep set
if QUICK:
    CASES = [c for c in CASES if (c[0], c[1]) in (("kuramoto", 100003), ("nbody", 40001), ("nbody", 5003), ("kuramoto_lowrank64", 4099))]
if ONLY == "leap":
    CASES = []
for model, n, meth, nsteps, tol, opts in CASES:
    lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", model + ".pop"))
    lo, hi = odist.shard_range(n, rank, world)
    pop = capi.Population(lib, capacity=hi - lo + 64, device=local)
    odist.connect(pop)
    for k, v in opts.items():
        pop.set_option(k, v)
    pop.set_seed(77)
    pop.init(hi - lo, replay=False, seed=77)
    t, h = 0.0, 0.01
    if meth == "rkf45":
        pop.set_tolerances(1e-9, 1e-9)
        pop.evolve_reset()
        for _ in range(nsteps):
            t, h = pop.evolve_apply("rkf45", t, 1.0, h)
    else:
        t = pop.run_fixed(meth, 0.0, 0.01, nsteps)
    pop.rhs(t)
    mine = pop.download()
    mfd = pop.meanfields()
    pop.sync()
    if rank == 0:
        solo = capi.Population(lib, capacity=n + 64, device=local)
        solo.set_seed(77)
        solo.init(n, replay=False, seed=77)
        ts, hs = 0.0, 0.01
        if meth == "rkf45":
            solo.set_tolerances(1e-9, 1e-9); solo.evolve_reset()
            for _ in range(nsteps):
                ts, hs = solo.evolve_apply("rkf45", ts, 1.0, hs)
        else:
            ts = solo.run_fixed(meth, 0.0, 0.01, nsteps)
        solo.rhs(ts)
        ref = solo.download(lo, hi - lo)
        dy = np.abs(mine["y"] - ref["y"]).max() / (np.abs(ref["y"]).max() + 1e-300)
        dp = (np.abs(mine["psi"] - ref["psi"]).max() / (np.abs(ref["psi"]).max() + 1e-300)) if lib.npsi else 0.0
        dm = (np.abs(mfd - solo.meanfields()) / np.maximum(1.0, np.abs(mfd))).max() if lib.nmfd else 0.0      # {vect} fields scale with POP_SIZE
        good = (dy <= tol) and (dp <= max(tol, 1e-12)) and (dm <= 1e-12) and np.array_equal(mine["id"], ref["id"]) and abs(t - ts) < 1e-15
        ok &= bool(good)
        print("%-8s %-5s %-22s N=%d world=%d  t=%.6g (solo %.6g)  max rel dy=%.2e  dpsi=%.2e  dmfd=%.2e  ids_equal=%s  -> %s" % (
            model, meth, opts or "", n, world, t, ts, dy, dp, dm, np.array_equal(mine["id"], ref["id"]), "OK" if good else "MISMATCH"), flush=True)
        solo.close()
    dist.barrier()
    pop.close()
# ---- tau-leaping sweeps on a sharded population: sharded sweep + redistribution (and the replicated form) against the solo run
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "vdp_sde.pop"))
for n, shard in (((200001, 1),) if QUICK else ((200001, 1), (200001, 0), (5, 1), (70001, 1))):
    lo, hi = odist.shard_range(n, rank, world)
    pop = capi.Population(lib, capacity=2 * (hi - lo) + 4096, device=local)
    odist.connect(pop)
    pop.set_option("shard_events", shard)
    pop.init(hi - lo, replay=False, seed=5)
    t = pop.run_fixed("fe", 0.0, 0.01, 20)
    solo = None
    if rank == 0:
        solo = capi.Population(lib, capacity=2 * n + 4096, device=local)
        solo.init(n, replay=False, seed=5)
        solo.run_fixed("fe", 0.0, 0.01, 20)
    good = True
    for k in range(4):
        pop.rhs(t)
        rows = pop.tauleap_apply(t, 2.0 if n > 100 else 8.0, k, as_array=True)
        mine = pop.download()
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine["id"], mine["y"]))
        if rank == 0:
            solo.rhs(t)
            srows = solo.tauleap_apply(t, 2.0 if n > 100 else 8.0, k, as_array=True)
            ref = solo.download()
            ids_all = np.concatenate([g[0] for g in gathered]); y_all = np.concatenate([g[1] for g in gathered])
            sizes = [len(g[0]) for g in gathered]
            ok_k = (rows.tobytes() == srows.tobytes() and np.array_equal(ids_all, ref["id"]) and np.array_equal(y_all, ref["y"])
                    and max(sizes) - min(sizes) <= 1 and pop.size()[1] == solo.size()[0] and pop.size()[2] == solo.size()[2])
            good &= bool(ok_k)
            print("tau-leap sweep %d N=%d world=%d %s: %d rows (solo %d), N -> %d (solo %d), shards %s -> %s" % (
                k, n, world, "sharded" if shard else "replicated", len(rows), len(srows), pop.size()[1], solo.size()[0], sizes, "OK" if ok_k else "MISMATCH"), flush=True)
    ok &= good
    if rank == 0:
        solo.close()
    dist.barrier()
    pop.close()
# ---- whole simulation with birth / death through the C host driver: every rank's stats.dat must equal the solo run's
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "vdp_sde.pop"))
for ssahmin, ps, res, T in ([(1.0, 2000, 21, 2.0)] if QUICK else [(1.0, 2000, 21, 2.0), (0.0, 60, 6, 1.0)]):
    pop = capi.Population(lib, capacity=4 * ps + 1024, device=local)
    odist.connect(pop)
    out = capi.odesolver(pop, (0, T), meth="fe", res=res, popsize=ps, h0=0.01, ssahmin=ssahmin, seed=123, replay_ic=False)
    ids_mine = out["traj"]["id"][out["traj"]["step"] == out["traj"]["step"].max()]
    gathered = [None] * world
    dist.all_gather_object(gathered, ids_mine)
    if rank == 0:
        solo = capi.Population(lib, capacity=4 * ps + 1024, device=local)
        ref = capi.odesolver(solo, (0, T), meth="fe", res=res, popsize=ps, h0=0.01, ssahmin=ssahmin, seed=123, replay_ic=False)
        sa, sb = out["stats"], ref["stats"]
        same = len(sa) == len(sb) and all(np.array_equal(sa[k], sb[k]) for k in ("n", "stop_flag", "event")) and np.allclose(sa["t"], sb["t"], rtol=1e-12)
        ids_ref = ref["traj"]["id"][ref["traj"]["step"] == ref["traj"]["step"].max()]
        ids_all = np.concatenate(gathered)
        good = bool(same and np.array_equal(ids_all, ids_ref) and max(len(g) for g in gathered) - min(len(g) for g in gathered) <= 1)
        ok &= good
        print("vdp_sde odesolver ssahmin=%g ps=%d world=%d: stats rows %d (solo %d) events %d final N %d shards %s -> %s" % (
            ssahmin, ps, world, len(sa), len(sb), int((sb["event"][:, 1] != 0).sum()) - ps, int(sb["n"][-1]), [len(g) for g in gathered], "OK" if good else "MISMATCH"), flush=True)
        solo.close()
    dist.barrier()
    pop.close()
# ---- a rank that asks for another number of steps: every rank gets an error from odx_sync instead of a hung mean-field barrier
lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "kuramoto.pop"))
lo, hi = odist.shard_range(100003, rank, world)
pop = capi.Population(lib, capacity=hi - lo + 64, device=local)
odist.connect(pop)
pop.init(hi - lo, replay=False, seed=77)
pop.run_fixed("dde", 0.0, 0.01, 5); pop.sync()
msg = ""
try:
    pop.run_fixed("dde", 0.05, 0.01, 3 if rank == 0 else 4)
    pop.sync()
except Exception as e:
    msg = str(e)
flags = [None] * world
dist.all_gather_object(flags, "disagree" in msg)
if rank == 0:
    good = all(flags)
    ok &= good
    print("step-count mismatch between ranks: %s -> %s" % (msg[:60] or "no error", "OK" if good else "MISMATCH"), flush=True)
dist.barrier()
pop.close()
if rank == 0:
    print("MGPU_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
