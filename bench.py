#!/usr/bin/env python3
"""bench.py — particle-steps/s of the population hot path (BASELINE.json metric) on N B200s.

Workload (configs[1]): Kuramoto mean-field model, N = 2^24 oscillators, classic fixed-step RK4 (`meth dde`), FP64.
A "step" is one RK4 step of the whole population (4 RHS evaluations = 4 fused stage kernels with one mean-field
reduction each).  Multi-GPU: particles sharded by contiguous ranges, total N fixed ("strong" scaling, as the config
names one N for 1/2/4/8 GPUs), one allreduce of the %M sums per stage.

  value      particle-steps/s with the population resident in HBM (device-timed, CUDA events on the launch stream)
  e2e        the same metric through the host-buffer C-ABI path: every step uploads (theta, omega) from pinned host
             memory, takes one RK4 step and reads the mean fields back
  roofline   algorithmic bytes (72 B per particle-step, SURVEY.md §8d / DESIGN.md) / measured step time vs the HBM peak
  cpu_baseline  the unchanged reference binary (oracle/_ref) on one host core, bounded sample

`--impl reference` times the reference's CPU path (oracle/_ref, else the oracle port) with the same JSON line.
"""
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec (RK4, FP64)"
N_TOTAL = 1 << 24
H_STEP = 0.05
ALG_BYTES_PER_PARTICLE_STEP = 72.0
NCU_DRAM_BYTES_PER_PARTICLE_STEP = 96.4    # measured, see roofline.traffic below
L2_BYTES = 132644864                       # cudaDeviceProp::l2CacheSize of the B200 (the library's variant thresholds use it)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of the GPUs in use during the timed region (all ranks' GPUs: in a
    strong-scaling run every stage waits for the slowest GPU, so one power-capped GPU sets the pace)"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, ngpus):
        self.ngpus, self.rows, self.proc = ngpus, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        per, mx, reasons = {}, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                idx = int(r[0])
                if idx >= self.ngpus:
                    continue
                per.setdefault(idx, []).append(float(r[1])); mx.append(float(r[2]))
                for k, nme in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        med = {i: sorted(v)[len(v) // 2] for i, v in per.items() if v}
        return {"sm_mhz": min(med.values()) if med else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": sum(len(v) for v in per.values()), "per_gpu_sm_mhz": [med[i] for i in sorted(med)]}


def cpu_reference_sample(n, nsteps, meth="dde"):
    """time the unchanged reference binary (1 core) on a bounded sample of the workload; returns (value, info)"""
    from oracle import refrun
    if refrun.have("kuramoto"):
        r = refrun.run("kuramoto", {"meth": meth, "h0": H_STEP, "res": 2, "ps": n, "hmin": 1e-9}, timespan=(0, H_STEP * nsteps),
                       read_traj=False, timeout=1800)
        ms = r.get("solver_ms")
        if ms:
            return n * nsteps / (ms * 1e-3), {"kind": "reference", "cores": 1,
                                             "sample": "Kuramoto N=%d, %d RK4 steps (meth dde, h0=%g, res=2), reference's own `solver` timer" % (n, nsteps, H_STEP)}
    try:
        from oracle import port
        v = port.time_kuramoto_rk4(n, nsteps, H_STEP)
        return v, {"kind": "port", "cores": 1, "sample": "oracle C restatement, Kuramoto N=%d, %d RK4 steps" % (n, nsteps)}
    except Exception as e:
        return None, {"kind": "port", "cores": 1, "sample": "unavailable: %s" % e}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, per = 1 << 18, 4
    for _ in range(args.warmup):
        cpu_reference_sample(1 << 14, 2)
    t0 = time.time()
    vals = []
    info = {}
    for _ in range(args.steps):
        v, info = cpu_reference_sample(n, per)
        vals.append(v)
    wall = time.time() - t0
    vals = [v for v in vals if v]
    value = sum(vals) / len(vals) if vals else None
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Kuramoto mean-field N=2^24 RK4 (dde) FP64; CPU arm runs a bounded sample per step", "sample_n": n,
                       "sample_rk4_steps": per},
            "cpu_baseline": dict(info, value=value, unit="particle-steps/s"),
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="odx")
    ap.add_argument("--n", type=int, default=N_TOTAL)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-peers", action="store_true", help="per-stage sums through NCCL allreduce instead of NVLink peer mailboxes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from odexp_b200 import capi

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = capi.ModelLib.from_pop(os.path.join(ROOT, "models", "kuramoto.pop"))
    n_total = args.n
    from odexp_b200.dist import shard_range
    lo, hi = shard_range(n_total, rank, world)
    n_local = hi - lo
    pop = capi.Population(lib, capacity=n_local, device=local)
    if world > 1:
        from odexp_b200 import dist as odist
        odist.connect(pop, use_nccl=True, use_peers=not args.no_peers)
    # synthetic population: theta0 = 2*PI*u, omega ~ N(2*PI/10, 0.1), keyed Philox on the device (ode_init_conditions)
    pop.init(n_local, t=0.0, seed=3141592, replay=False)
    pop.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        pop.sync()

    t = 0.0
    t = pop.run_fixed("dde", t, H_STEP, args.warmup)
    barrier()
    sampler = ClockSampler(world)
    if rank == 0:
        sampler.start()
    l0 = pop.launch_count()
    pop.event_record(0)
    t = pop.run_fixed("dde", t, H_STEP, args.steps)
    pop.event_record(1)
    ms = pop.event_elapsed_ms(0, 1)
    barrier()
    launches = pop.launch_count() - l0
    # the timed region lasts tens of milliseconds: too short for more than a few nvidia-smi samples.  Keep the sampler running
    # over an UNTIMED continuation of the same workload (same kernel, same population) so that the clocks of every GPU under
    # this load are on record: in a strong-scaling run one power-capped GPU sets the pace of all.
    t_samples = len(sampler.rows) if rank == 0 else 0
    if world > 1:                                 # device time of the timed region: MAX over ranks
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    # every rank MUST run the same number of steps (the step kernel exchanges the mean-field sums with its peers every stage: a
    # rank with one step more would wait for ever), so the count is derived from the reduced time and then taken from rank 0
    extra = max(0, min(4000, int(400.0 / max(ms / args.steps, 1e-3))))          # ~0.4 s
    if world > 1:
        from odexp_b200 import dist as odist
        extra = int(odist.agreed(extra, src=0))
    t = pop.run_fixed("dde", t, H_STEP, extra)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["samples_in_timed_region"] = t_samples
        clocks["note"] = "sampling continued over %d untimed steps of the same workload after the timed region" % extra
    mfd = pop.meanfields()
    value = n_total * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    ne2e = args.e2e_steps
    st = pop.download()                       # current state as host arrays (particle-major)
    y_pin = torch.from_numpy(st["y"].copy()).pin_memory()
    e_pin = torch.from_numpy(st["expr"].copy()).pin_memory()
    del st
    for _ in range(2):
        pop.upload_ptr(n_local, y_pin.data_ptr(), e_pin.data_ptr()); pop.run_fixed("dde", 0.0, H_STEP, 1); pop.meanfields()
    barrier()
    w0 = time.perf_counter()
    for _ in range(ne2e):
        pop.upload_ptr(n_local, y_pin.data_ptr(), e_pin.data_ptr())
        pop.run_fixed("dde", 0.0, H_STEP, 1)
        pop.meanfields()                       # D2H of the step's result (mean fields) + sync
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
    e2e_value = n_total * ne2e / e2e_s
    # the same path used the way odesolver uses it: ONE upload of the host state, args.steps RK4 steps, ONE download of the final
    # state (information only; the contract's e2e above re-uploads the inputs for every step)
    barrier()
    w0 = time.perf_counter()
    pop.upload_ptr(n_local, y_pin.data_ptr(), e_pin.data_ptr())
    pop.run_fixed("dde", 0.0, H_STEP, args.steps)
    fin = pop.download()
    barrier()
    sess_s = time.perf_counter() - w0
    if world > 1:
        ts_ = torch.tensor([sess_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
        sess_s = float(ts_.item())
    sess_bytes = (int(n_local * 16), int(sum(v.nbytes for v in fin.values())))
    del fin

    if rank == 0:
        peak, how = peaks()
        achieved = ALG_BYTES_PER_PARTICLE_STEP * n_local * args.steps / (ms * 1e-3) / 1e9     # per GPU
        line = {"metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "Kuramoto mean-field (models/kuramoto.pop), N=%d oscillators, classic RK4 (meth dde) h=%g, FP64" % (n_total, H_STEP),
                           "n_particles": n_total, "per_gpu_particles": n_local, "l2": ("state larger than L2 (%.0f MB per GPU vs 126 MB): every stage streams it from HBM" % (n_local * 32 / 1e6)) if n_local * 32 > 126e6
                                 else ("state fits in L2 (%.0f MB per GPU): consecutive steps of one simulation reuse the resident population by construction, "
                                       "no flush between steps (a flush would not be part of the workload)" % (n_local * 32 / 1e6)),
                           "stage_sweeps_in_timed_region": int(4 * args.steps),
                           "parallelism": "particle sharding x%d, %%M sums per stage through %s" % (world, "NCCL allreduce" if args.no_peers else "in-kernel NVLink peer mailboxes"), "mean_fields_after": [float(x) for x in mfd]},
                "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(n_local * 16), "d2h_bytes_per_step": int(8 * lib.nmfd),
                        "steps": ne2e, "what": "odx_pop_upload(host theta, omega) + odx_run_fixed(dde, 1 step) + odx_meanfields per step"},
                "e2e_session": {"value": n_total * args.steps / sess_s, "unit": "particle-steps/s", "steps": args.steps, "h2d_bytes_total": sess_bytes[0],
                                "d2h_bytes_total": sess_bytes[1], "what": "one odx_pop_upload + odx_run_fixed(dde, %d steps) + one odx_pop_download (host wall clock)" % args.steps},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             # dram__bytes_read.sum + dram__bytes_write.sum of k_chain_persist from the ncu --set full capture
                             # profiles/r1_k_chain_persist_lean_kuramoto_n16M.txt: 32.34 GB for 20 steps at N=2^24 = 96.4 B per
                             # particle-step (lean RK4: 112 B by construction, minus what the alternating sweep direction finds in L2)
                             "traffic": (NCU_DRAM_BYTES_PER_PARTICLE_STEP * n_local * args.steps) if (world == 1 and n_total == N_TOTAL) else None,
                             "traffic_unit": "bytes per launch (one cooperative launch = all timed steps)",
                             "peak_source": how, "algorithmic_bytes_per_particle_step": ALG_BYTES_PER_PARTICLE_STEP,
                             "algorithmic_bytes_per_launch": ALG_BYTES_PER_PARTICLE_STEP * n_local * args.steps,
                             "kernel": "k_chain_persist (one cooperative launch runs all timed RK4 steps: 4 stages per step, in-kernel mean-field exchange; "
                                       + ("TMA ring + lean RK4 stages" if n_local * 32 > 1.5 * L2_BYTES else "TMA ring + stored stages" if n_local * 32 > 0.9 * L2_BYTES
                                          else "plain loads + stored stages, L2-resident shard") + ")"}}
        if world == 1 and not args.no_cpu:
            v, info = cpu_reference_sample(1 << 20, 16)
            line["cpu_baseline"] = dict(info, value=v, unit="particle-steps/s")
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    pop.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
