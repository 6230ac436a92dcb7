#!/usr/bin/env python3
"""bench.py — particle-steps/s of the population hot path (BASELINE.json metric) on N B200s.

Workload (configs[1]): Kuramoto mean-field model, N = 2^24 oscillators, classic fixed-step RK4 (`meth dde`), FP64.
A "step" is one RK4 step of the whole population (4 RHS evaluations = 4 fused stage sweeps with one mean-field
reduction each).  Multi-GPU: particles sharded by contiguous ranges, total N fixed ("strong" scaling, as the config
names one N for 1/2/4/8 GPUs), one exchange of the %M sums per stage inside the step kernel.

  value         particle-steps/s with the population resident in HBM (device-timed, CUDA events on the launch stream)
  e2e           the same metric through the host-buffer C-ABI path: every step uploads (theta, omega) from pinned host
                memory, takes one RK4 step and reads the mean fields back
  roofline      algorithmic bytes (72 B per particle-step, SURVEY.md 8d / DESIGN.md) / measured step time vs the HBM peak
  roofline_fp64 FP64 instructions the step kernel executes (ncu count, profiles/) / measured step time vs the measured
                DFMA issue peak (profiles/r1_fp64_peak.txt)
  parity        GPU vs the unchanged reference binary on the same replayed initial conditions, in this run
  configs       the other BASELINE configurations (N-body RKF45, low-rank rank 64, vdP SDE + tau-leaping), each with
                its own roofline fraction
  cpu_baseline  the unchanged reference binary (oracle/_ref) on one host core, bounded sample

Timed region (immune to host-side launch skew between ranks): barrier, then with NO host synchronisation in between
  run_fixed(1)  -> event 0 -> run_fixed(K) -> event 1 -> run_fixed(continuation) -> event 2
The single step in front absorbs any difference in launch time inside the kernel's first mean-field exchange; events
0/1 bracket exactly the K timed steps on every rank, the MAX over ranks is reported (per-rank values are on the line).
The continuation (same kernel, same population) is timed as well: `ms_per_step_sustained`; clocks are sampled through
NVML over both.

`--impl reference` times the reference's CPU path (oracle/_ref, else the oracle port) with the same JSON line.
"""
import argparse, json, os, signal, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec (RK4, FP64)"
N_TOTAL = 1 << 24
H_STEP = 0.05
ALG_BYTES_PER_PARTICLE_STEP = 72.0
# measured with ncu --set full on k_chain_persist at N = 2^24 (profiles/, per particle-step): DRAM bytes and executed FP64
# thread-instructions.  Constants because a number taken under a profiler must not be produced by the bench itself.
NCU = {"ring-lean": {"dram_bytes": 99.3, "fp64_inst": 196.7, "file": "profiles/r2r_k_chain_persist_lean_kuramoto_n16M.txt (round 2: table sincos; "
                     "8.33 GB of DRAM traffic and 515.6 M FP64 warp instructions for 5 steps of 2^24 particles)"}}
FP64_PEAK_INST = 18.47e12                  # DFMA/s, reg,reg,const form, 64 warps/SM: profiles/r1_fp64_peak.txt (scripts/ubench/fp64_issue.cu)
FP64_PEAK_FLOPS = 2 * FP64_PEAK_INST
L2_BYTES = 132644864                       # cudaDeviceProp::l2CacheSize of the B200 (the library's variant thresholds use it)
MODELS = os.path.join(ROOT, "models")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(n_total, world):
    n_local = (n_total + world - 1) // world
    big = n_local * 32 > 126e6
    return {"workload": "Kuramoto mean-field (models/kuramoto.pop = the reference's examples/kuramoto.pop), N=%d oscillators, "
                        "classic RK4 (meth dde) h=%g, FP64" % (n_total, H_STEP),
            "n_particles": n_total, "method": "dde", "h": H_STEP,
            "l2": ("inputs larger than L2: %.0f MB of state per GPU vs 126 MB, every stage streams it from HBM" % (n_local * 32 / 1e6)) if big
                  else ("state fits in L2 (%.0f MB per GPU): consecutive steps of one simulation reuse the resident population by "
                        "construction, no flush between steps (a flush would not be part of the workload)" % (n_local * 32 / 1e6))}


class ClockSampler:
    """SM clocks / throttle reasons / power of the GPUs in use, sampled through NVML from a thread of this process (no fork
    between the barrier and the launches; ~1 ms period so that even a 7 ms timed region holds samples).  Falls back to
    `nvidia-smi -lms` when pynvml is unavailable."""
    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, ngpus):
        self.ngpus, self.rows, self.stop_ev, self.th, self.proc, self.how = ngpus, [], threading.Event(), None, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.hs = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in range(self.ngpus)]
            self.max_mhz = max(float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)) for h in self.hs)
            self.how = "nvml"
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.how = None
        try:
            q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        k = 0
        while not self.stop_ev.is_set():
            for i, h in enumerate(self.hs):
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                    pw = float(nv.nvmlDeviceGetPowerUsage(h)) / 1e3 if (k & 15) == 0 else None
                    self.rows.append((time.perf_counter(), i, mhz, rs, pw))
                except Exception:
                    pass
            k += 1
            time.sleep(0.0005)

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                rs = sum(bit for bit, nme in zip((0x8, 0x40, 0x20, 0x4), r[4:8]) if nme.lower().startswith("active"))
                self.max_mhz = max(self.max_mhz or 0.0, float(r[2]))
                self.rows.append((time.perf_counter(), int(r[0]), float(r[1]), rs, float(r[3])))
            except Exception:
                continue

    def stop(self, windows):
        """windows: {name: (t0, t1)} host-clock windows; returns the clocks record (medians over every sample from the first
        window's start to the last window's end, per-window sample counts)"""
        self.stop_ev.set()
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.th:
            self.th.join(timeout=2)
        if not self.how:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"]}
        lo = min(w[0] for w in windows.values()); hi = max(w[1] for w in windows.values())
        per, reasons, pw = {}, set(), []
        for ts, i, mhz, rs, p in self.rows:
            if i >= self.ngpus or ts < lo or ts > hi:
                continue
            per.setdefault(i, []).append(mhz)
            for bit, nme in self.NAMES.items():
                if rs & bit:
                    reasons.add(nme)
            if p is not None:
                pw.append(p)
        med = {i: sorted(v)[len(v) // 2] for i, v in per.items() if v}
        out = {"sm_mhz": min(med.values()) if med else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
               "samples": sum(len(v) for v in per.values()), "per_gpu_sm_mhz": [med[i] for i in sorted(med)],
               "power_w_max": max(pw) if pw else None, "source": self.how}
        for name, (a, b) in windows.items():
            out["samples_in_" + name] = sum(1 for ts, i, *_ in self.rows if i < self.ngpus and a <= ts <= b)
        return out


def cpu_reference_sample(n, nsteps, meth="dde", read_traj=False):
    """time the unchanged reference binary (1 core) on a bounded sample of the workload; returns (value, info, run)"""
    from oracle import refrun
    if refrun.have("kuramoto"):
        dims = dict(nvar=1, naux=2, npsi=0, nmfd=3, nexpr=1) if read_traj else None
        r = refrun.run("kuramoto", {"meth": meth, "h0": H_STEP, "res": 2, "ps": n, "hmin": 1e-9}, timespan=(0, H_STEP * nsteps),
                       read_traj=read_traj, dims=dims, timeout=1800)
        ms = r.get("solver_ms")
        if ms:
            return n * nsteps / (ms * 1e-3), {"kind": "reference", "cores": 1,
                                             "sample": "Kuramoto N=%d, %d RK4 steps (meth dde, h0=%g, res=2), the unchanged reference binary "
                                                       "(oracle/_ref), its own `solver` timer; the reference is single-threaded" % (n, nsteps, H_STEP)}, r
    try:
        from oracle import port
        v = port.time_kuramoto_rk4(n, nsteps, H_STEP)
        return v, {"kind": "port", "cores": 1, "sample": "oracle C restatement, Kuramoto N=%d, %d RK4 steps" % (n, nsteps)}, None
    except Exception as e:
        return None, {"kind": "port", "cores": 1, "sample": "unavailable: %s" % e}, None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    n, per = 1 << 18, 4
    for _ in range(args.warmup):
        cpu_reference_sample(1 << 14, 2)
    t0 = time.time()
    vals = []
    info = {}
    for _ in range(args.steps):
        v, info, _ = cpu_reference_sample(n, per)
        vals.append(v)
    wall = time.time() - t0
    vals = [v for v in vals if v]
    value = sum(vals) / len(vals) if vals else None
    info = dict(info, sample=info.get("sample", "") + "; every bench step = one such run (the per-particle rate of the reference falls with N: "
                "1.5e6 at 2^16, 1.0e6 at 2^20 (BASELINE.md 2), so the sample is generous to the CPU)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.n, world),
            "cpu_baseline": dict(info, value=value, unit="particle-steps/s"),
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# the other BASELINE configurations (device-timed on the library's stream, every rank runs the same calls)
def _connect(pop, world, no_peers):
    if world > 1:
        from odexp_b200 import dist as odist
        odist.connect(pop, use_nccl=True, use_peers=not no_peers)


def _max_over_ranks(x, world):
    if world == 1:
        return x
    import torch, torch.distributed as dist
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def config_nbody(capi, odist, rank, world, local, no_peers, peak_hbm):
    """configs[3]: N-body N = 2^17, O(N^2) %C, RKF45 with the global error norm (allgather of the coupled columns + MAX allreduce)"""
    n = 1 << 17
    lib = capi.ModelLib.from_pop(os.path.join(MODELS, "nbody.pop"))
    lo, hi = odist.shard_range(n, rank, world)
    pop = capi.Population(lib, capacity=hi - lo, device=local)
    _connect(pop, world, no_peers)
    pop.init(hi - lo, replay=False, seed=11)
    pop.rhs(0.0); pop.sync()
    reps = 3
    pop.event_record(0)
    for _ in range(reps):
        pop.rhs(0.0)
    pop.event_record(1)
    ms_eval = _max_over_ranks(pop.event_elapsed_ms(0, 1) / reps, world)
    ms_ordered = None
    if world == 1:                                        # the ordered kernel (k_pair: the reference's summation order) beside the symmetric one
        pop.set_option("pair_sym", 0)
        pop.rhs(0.0); pop.sync()
        pop.event_record(0)
        pop.rhs(0.0)
        pop.event_record(1)
        ms_ordered = pop.event_elapsed_ms(0, 1)
        pop.set_option("pair_sym", -1)
        pop.rhs(0.0); pop.sync()
    pop.set_tolerances(1e-8, 1e-8)
    pop.evolve_reset()
    t, h = 0.0, 1e-3
    t, h = pop.evolve_apply("rkf45", t, 1.0, h)           # first step evaluates dydt_in; the timed ones reuse dydt_out like every later step
    pop.sync()
    nacc = 3
    c0 = pop.evolve_stats()
    pop.event_record(0)
    for _ in range(nacc):
        t, h = pop.evolve_apply("rkf45", t, 1.0, h)
    pop.event_record(1)
    ms_step = _max_over_ranks(pop.event_elapsed_ms(0, 1) / nacc, world)
    c1 = pop.evolve_stats()
    pop.close()
    pairs = float(n) * n
    flops = 20.0 * pairs / (ms_eval * 1e-3) / world         # per GPU
    sym = True                                               # 128 tiles of 1024 bodies: the symmetric kernel on any number of ranks
    inst = 9.0 if sym else 15.0                              # executed FP64 instructions per ordered pair
    return {"workload": "models/nbody.pop (from examples/Nbody.pop), N=%d bodies, O(N^2) %%C gravity" % n, "n_particles": n,
            "ms_per_rhs_eval": ms_eval, "ms_per_rhs_eval_ordered_kernel": ms_ordered, "pairs_per_s": pairs / (ms_eval * 1e-3),
            "rkf45": {"ms_per_accepted_step": ms_step, "particle_steps_per_s": n / (ms_step * 1e-3), "accepted": nacc,
                      "rejected_attempts": int(c1[1] - c0[1]), "rhs_evals": int(c1[2] - c0[2]), "abstol": 1e-8, "reltol": 1e-8},
            "roofline": {"bound": "fp64", "achieved": flops / 1e12, "peak": FP64_PEAK_FLOPS / 1e12, "unit": "TFLOP/s", "frac": flops / FP64_PEAK_FLOPS,
                         "algorithmic_flops_per_pair": 20, "peak_source": "measured DFMA issue rate x 2 (profiles/r1_fp64_peak.txt); nominal 37.2",
                         "kernel": "k_pair_sym (every unordered pair once, added to both particles; 20 flop are credited per ORDERED pair, "
                                   "the convention of SURVEY 8(d); sharded: allgather of the coupled columns, half of every cross-rank tile pair per rank, "
                                   "the sums for the peers' bodies sent to their owners by ncclSend / ncclRecv)" if sym else "k_pair (ordered)",
                         "fp64_pipe_frac": inst * pairs / (ms_eval * 1e-3) / world / (FP64_PEAK_FLOPS / 2.0),
                         "executed_fp64_inst_per_ordered_pair": inst}}


def config_lowrank(capi, odist, rank, world, local, no_peers, peak_hbm):
    """configs[2]: kuramoto_lowrank, N = 2^20, rank 64 (allreduce of the r x (p+1) moments)"""
    n = 1 << 20
    lib = capi.ModelLib.from_pop(os.path.join(MODELS, "kuramoto_lowrank64.pop"))
    lo, hi = odist.shard_range(n, rank, world)
    pop = capi.Population(lib, capacity=hi - lo, device=local)
    _connect(pop, world, no_peers)
    pop.init(hi - lo, replay=False, seed=12)
    pop.rhs(0.0); pop.sync()
    reps = 5
    pop.event_record(0)
    for _ in range(reps):
        pop.rhs(0.0)
    pop.event_record(1)
    ms_eval = _max_over_ranks(pop.event_elapsed_ms(0, 1) / reps, world)
    p_sel = pop.lowrank_order()[0]
    t = pop.run_fixed("dde", 0.0, H_STEP, 2); pop.sync()
    pop.event_record(0); t = pop.run_fixed("dde", t, H_STEP, 5); pop.event_record(1)
    ms_step = _max_over_ranks(pop.event_elapsed_ms(0, 1) / 5, world)
    pop.close()
    nl = hi - lo
    gbs = (16.0 * nl * 64 + 16.0 * nl) / (ms_eval * 1e-3) / 1e9
    return {"workload": "models/kuramoto_lowrank64.pop (examples/kuramoto_lowrank.pop with rank 64), N=%d" % n, "n_particles": n,
            "ms_per_rhs_eval": ms_eval, "selected_order_p": p_sel, "rk4_ms_per_step": ms_step, "rk4_particle_steps_per_s": n / (ms_step * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                         "algorithmic_bytes_per_eval_per_gpu": 16.0 * nl * 64 + 16.0 * nl, "kernel": "k_lr_moments_tma + k_lr_apply_tma (whole evaluation timed)"}}


def config_sde(capi, odist, rank, world, local, no_peers, peak_hbm):
    """configs[4]: van der Pol SDE population, N = 2^22, Euler-Maruyama with keyed Philox noise + tau-leaping birth / death"""
    n = 1 << 22
    lib = capi.ModelLib.from_pop(os.path.join(MODELS, "vdp_sde.pop"))
    lo, hi = odist.shard_range(n, rank, world)
    pop = capi.Population(lib, capacity=2 * (hi - lo) + 8192, device=local)
    _connect(pop, world, no_peers)
    pop.init(hi - lo, replay=False, seed=8)
    t = pop.run_fixed("fe", 0.0, 0.001, 20); pop.sync()
    ns = 200
    pop.event_record(0); t = pop.run_fixed("fe", t, 0.001, ns); pop.event_record(1)
    ms_step = _max_over_ranks(pop.event_elapsed_ms(0, 1) / ns, world)
    pop.rhs(t); pop.sync()
    sweeps = []
    for k in range(8):                  # as in a simulation: a run of Euler-Maruyama steps, one RHS evaluation, one sweep
        t = pop.run_fixed("fe", t, 0.001, 50)
        pop.rhs(t)
        n0 = pop.size()[1]
        pop.event_record(0)
        rows = pop.tauleap_apply(t, 1.0, k, as_array=True)
        pop.event_record(1)
        sweeps.append({"n_before": int(n0), "events": int(len(rows)), "ms": _max_over_ranks(pop.event_elapsed_ms(0, 1), world)})
    pop.sync()
    pop.close()
    nl = hi - lo
    gbs = 48.0 * nl / (ms_step * 1e-3) / 1e9
    best = sorted(s["ms"] for s in sweeps[4:])[2]          # median of the last four (the first ones include allocations and kernel loading)
    return {"workload": "models/vdp_sde.pop, N=%d, Euler-Maruyama (meth fe) h=0.001, keyed Philox noise, tau-leaping sweeps" % n, "n_particles": n,
            "em_ms_per_step": ms_step, "em_particle_steps_per_s": n / (ms_step * 1e-3), "tauleap_sweeps": sweeps, "tauleap_ms_steady": best,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                         "algorithmic_bytes_per_particle_step": 48, "kernel": "k_chain_persist (Euler-Maruyama)"},
            "tauleap_roofline": {"bound": "hbm", "algorithmic_bytes_per_particle": 72, "what": "read 2 state + 2 rate columns + id, write the compacted 2 state "
                                 "+ 2 rate columns + id (36 + 36 B); one GPU: k_leap_sweep (one cooperative launch, device-side generation loop) + "
                                 "k_leap_compact, event rows copied while the compaction runs (profiles/r2s_tauleap_fused_sweep.txt); sharded: "
                                 "per-generation kernels + count allgathers + grouped send / recv rebalancing; once per leap interval",
                                 "frac": 72.0 * n / (best * 1e-3) / 1e9 / peak_hbm / world}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="odx")
    ap.add_argument("--n", type=int, default=N_TOTAL)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE configurations")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-peers", action="store_true", help="per-stage sums through NCCL allreduce instead of NVLink peer mailboxes")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (odx_set_option), e.g. lean=0")
    ap.add_argument("--skew-ms", type=float, default=0.0, help="diagnostic: rank 0 sleeps this long before enqueueing the timed launches")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from odexp_b200 import capi
    from odexp_b200 import dist as odist

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = capi.ModelLib.from_pop(os.path.join(MODELS, "kuramoto.pop"))
    n_total = args.n
    lo, hi = odist.shard_range(n_total, rank, world)
    n_local = hi - lo
    pop = capi.Population(lib, capacity=n_local, device=local)
    _connect(pop, world, args.no_peers)
    for kv in args.opt:
        k, v = kv.split("=")
        pop.set_option(k, float(v))
    # synthetic population: theta0 = 2*PI*u, omega ~ N(2*PI/10, 0.1), keyed Philox on the device (ode_init_conditions)
    pop.init(n_local, t=0.0, seed=3141592, replay=False)
    pop.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        pop.sync()

    sampler = ClockSampler(world)
    if rank == 0:
        sampler.start()                                 # before the barrier: nothing is forked or initialised next to the timed launches
    t = 0.0
    # warm-up, device-timed: its rate fixes the length of the continuation BEFORE the timed region, identically on every rank
    t = pop.run_fixed("dde", t, H_STEP, args.warmup)   # the warm-up proper: kernel loading, first exchanges, peer agreement
    pop.event_record(3)
    t = pop.run_fixed("dde", t, H_STEP, max(args.warmup, 5))
    pop.event_record(4)
    warm_ms = _max_over_ranks(pop.event_elapsed_ms(3, 4) / max(args.warmup, 5), world)      # rate of a warm launch, not of the first one
    extra = max(20, min(4000, int(400.0 / max(warm_ms, 1e-3))))          # ~0.4 s of the same workload
    extra = int(odist.agreed(extra, src=0))
    barrier()
    if args.skew_ms > 0 and rank == 0:
        time.sleep(args.skew_ms * 1e-3)
    l0 = pop.launch_count()
    h_a = time.perf_counter()
    t = pop.run_fixed("dde", t, H_STEP, 1)              # absorbs launch skew between ranks inside its first mean-field exchange
    pop.event_record(0)
    t = pop.run_fixed("dde", t, H_STEP, args.steps)
    pop.event_record(1)
    t = pop.run_fixed("dde", t, H_STEP, extra)
    pop.event_record(2)
    pop.event_elapsed_ms(0, 0); h_0 = time.perf_counter()      # host clock when event 0 / 1 / 2 completed (for the clock windows only)
    ms_mine = pop.event_elapsed_ms(0, 1); h_1 = time.perf_counter()
    ms_sus_mine = pop.event_elapsed_ms(1, 2); h_2 = time.perf_counter()
    launches = (pop.launch_count() - l0) - 2            # kernels between event 0 and event 1: one cooperative launch runs all timed steps
    barrier()
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (ms_mine, ms_sus_mine))
        per_rank_ms = [g[0] for g in gathered]; per_rank_sus = [g[1] for g in gathered]
    else:
        per_rank_ms, per_rank_sus = [ms_mine], [ms_sus_mine]
    ms = max(per_rank_ms); ms_sus = max(per_rank_sus)
    clocks = sampler.stop({"timed_region": (max(h_a, h_0 - 1e-3), h_1), "continuation": (h_1, h_2)}) if rank == 0 else None
    if clocks is not None:
        clocks["note"] = "NVML samples over the timed region and the timed continuation (%d more steps of the same workload)" % extra
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
    e2e_s = _max_over_ranks(time.perf_counter() - w0, world)
    e2e_value = n_total * ne2e / e2e_s
    # the same path used the way odesolver uses it: ONE upload of the host state, args.steps RK4 steps, ONE download of the final
    # state (information only; the contract's e2e above re-uploads the inputs for every step)
    barrier()
    w0 = time.perf_counter()
    pop.upload_ptr(n_local, y_pin.data_ptr(), e_pin.data_ptr())
    pop.run_fixed("dde", 0.0, H_STEP, args.steps)
    pop.sync()
    w1 = time.perf_counter()
    fin = pop.download()
    w2 = time.perf_counter()
    barrier()
    sess_s = _max_over_ranks(time.perf_counter() - w0, world)
    sess_bytes = (int(n_local * 16), int(sum(v.nbytes for v in fin.values())))
    dl_gbs = sess_bytes[1] / (w2 - w1) / 1e9
    w3 = time.perf_counter()
    fin = pop.download(out=fin)                 # the same download into buffers whose pages exist already
    dl_gbs_warm = sess_bytes[1] / (time.perf_counter() - w3) / 1e9
    del fin, y_pin, e_pin

    peak, how = peaks()
    wset = n_local * 32.0
    variant = "ring-lean" if wset > 1.5 * L2_BYTES else "ring-stored" if wset > 0.9 * L2_BYTES else "plain-stored"
    line = None
    if rank == 0:
        achieved = ALG_BYTES_PER_PARTICLE_STEP * n_local * args.steps / (ms * 1e-3) / 1e9     # per GPU
        ncu = NCU.get(variant) if not args.opt else None
        cfg = workload_config(n_total, world)
        line = {"metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "ms_per_step_sustained": ms_sus / extra, "sustained_steps": extra,
                "per_rank_ms_per_step": [x / args.steps for x in per_rank_ms],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "run": {"per_gpu_particles": n_local, "stage_sweeps_in_timed_region": int(4 * args.steps), "step_kernel_variant": variant,
                        "parallelism": "particle sharding x%d, %%M sums per stage through %s" % (world, "NCCL allreduce" if args.no_peers else "in-kernel NVLink peer mailboxes"),
                        "mean_fields_after": [float(x) for x in mfd], "options": args.opt,
                        "timing": "barrier; run_fixed(1); event0; run_fixed(%d); event1; run_fixed(%d); event2 -- no host sync in between, MAX over ranks" % (args.steps, extra)},
                "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(n_local * 16), "d2h_bytes_per_step": int(8 * lib.nmfd),
                        "steps": ne2e, "what": "odx_pop_upload(host theta, omega) + odx_run_fixed(dde, 1 step) + odx_meanfields per step"},
                "e2e_session": {"value": n_total * args.steps / sess_s, "unit": "particle-steps/s", "steps": args.steps, "h2d_bytes_total": sess_bytes[0],
                                "d2h_bytes_total": sess_bytes[1], "download_gb_per_s_rank0": dl_gbs, "download_gb_per_s_rank0_touched_buffers": dl_gbs_warm, "what": "one odx_pop_upload + odx_run_fixed(dde, %d steps) + one odx_pop_download (host wall clock)" % args.steps},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "frac_sustained": ALG_BYTES_PER_PARTICLE_STEP * n_local * extra / (ms_sus * 1e-3) / 1e9 / peak,
                             # dram__bytes_read.sum + dram__bytes_write.sum of k_chain_persist from the ncu --set full capture
                             "traffic": (ncu["dram_bytes"] * n_local * args.steps) if (ncu and n_total == N_TOTAL) else None,
                             "traffic_source": ncu["file"] if ncu else None,
                             "traffic_unit": "bytes per launch (one cooperative launch = all timed steps)",
                             "peak_source": how, "algorithmic_bytes_per_particle_step": ALG_BYTES_PER_PARTICLE_STEP,
                             "algorithmic_bytes_per_launch": ALG_BYTES_PER_PARTICLE_STEP * n_local * args.steps,
                             "kernel": "k_chain_persist (one cooperative launch runs all timed RK4 steps: 4 stages per step, in-kernel mean-field exchange; "
                                       + {"ring-lean": "TMA ring + lean RK4 stages", "ring-stored": "TMA ring + stored stages",
                                          "plain-stored": "plain loads + stored stages, L2-resident shard"}[variant] + ")"}}
        if ncu and n_total == N_TOTAL:
            inst = ncu["fp64_inst"] * n_local * args.steps / (ms * 1e-3)
            line["roofline_fp64"] = {"bound": "fp64", "achieved": inst / 1e12, "peak": FP64_PEAK_INST / 1e12, "unit": "T FP64 inst/s", "frac": inst / FP64_PEAK_INST,
                                     "executed_fp64_inst_per_particle_step": ncu["fp64_inst"], "source": ncu["file"],
                                     "peak_source": "measured DFMA issue rate (profiles/r1_fp64_peak.txt); a DFMA holds the issue port for two cycles, "
                                                    "so this pipe fraction is also the share of issue slots FP64 work takes",
                                     "binding": "max(hbm, fp64): see DESIGN.md 3.1"}
    pop.close()

    # ---- the headline is complete: from here on a hang or failure in a secondary configuration must not lose it
    def emit_and_exit(note=None):
        if rank == 0 and line is not None:
            if note:
                line.setdefault("configs", {})["note"] = note
            line.setdefault("cpu_baseline", None)
            print(json.dumps(line), flush=True)
        os._exit(0)

    def on_alarm(signum, frame):
        emit_and_exit("secondary configurations timed out; headline unaffected")
    signal.signal(signal.SIGALRM, on_alarm)

    if not args.no_configs:
        signal.alarm(240)
        configs = {}
        failed = None
        for name, fn in (("nbody_rkf45_n131072", config_nbody), ("lowrank64_n1048576", config_lowrank), ("vdp_sde_n4194304", config_sde)):
            try:
                configs[name] = fn(capi, odist, rank, world, local, args.no_peers, peak)
            except Exception as e:      # the other ranks may be inside a collective: do not join any further one
                failed = "%s: %s" % (name, e)
                break
        signal.alarm(0)
        if rank == 0:
            line["configs"] = configs
        if failed:
            signal.alarm(60)            # a rank that failed alone leaves the others inside a collective: they leave through their own alarm
            emit_and_exit("stopped at " + failed)

    # ---- CPU baseline + same-run parity (rank 0, single GPU): the reference and the GPU integrate the same replayed population
    if rank == 0:
        if world == 1 and not args.no_cpu:
            n_par, k_par = 1 << 20, 16
            v, info, ref = cpu_reference_sample(n_par, k_par, read_traj=True)
            line["cpu_baseline"] = dict(info, value=v, unit="particle-steps/s")
            try:
                if ref is not None and "traj" in ref:
                    pp = capi.Population(lib, capacity=n_par, device=local)
                    # srand(3141592) + the reference's serial draw order for the ICs, then the same odesolver loop (res 2, meth dde)
                    out = capi.odesolver(pp, (0, H_STEP * k_par), meth="dde", h0=H_STEP, hmin=1e-9, res=2, popsize=n_par, replay_ic=True, traj_stride=1)
                    pp.close()
                    tr = ref["traj"]; last = tr[tr["step"] == tr["step"].max()]
                    gt = out["traj"]; glast = gt[gt["step"] == gt["step"].max()]
                    th_ref = last["cols"][:, 0]
                    rel = float(np.max(np.abs(glast["cols"][:, 0] - th_ref) / np.maximum(np.abs(th_ref), 1e-300)))
                    dm = float(np.max(np.abs(out["stats"]["mfd"] - ref["stats"]["mfd"])))
                    got = {"id": glast["id"]}
                    line["parity"] = {"n": n_par, "rk4_steps": k_par, "max_rel_state": rel, "max_abs_mfd": dm,
                                      "ids_equal": bool(np.array_equal(got["id"], last["id"])), "tolerance": {"state_rel": 1e-10, "mfd_abs": 1e-12},
                                      "ok": bool(rel <= 1e-10 and dm <= 1e-12),
                                      "what": "theta(t1) and the mean fields of the GPU run vs the unchanged reference binary, same replayed libc rand() ICs"}
            except Exception as e:
                line["parity"] = {"error": str(e)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
