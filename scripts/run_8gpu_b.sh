#!/bin/bash
# second 8-GPU measurement of the round (development tool): exchange cost and the bench line after the 16-byte remote polls
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29751 scripts/sync_cost.py 2>&1 | grep "^world" > gpurun_out/r2zz_sync_cost_8gpu.txt
cat gpurun_out/r2zz_sync_cost_8gpu.txt
$T --nproc-per-node 8 --master-port 29752 bench.py --gpus 8 --steps 20 --warmup 5 2>&1 | grep "^{" > gpurun_out/r2zz_bench_n8.json
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2zz_bench_n8.json").read()); c = d["configs"]
print(d["n_gpus"], "ms/step", d["ms_per_step"], "sustained", d.get("ms_per_step_sustained"), d["clocks"]["sm_mhz"], d["clocks"]["samples"], "nbody", c["nbody_rkf45_n131072"]["ms_per_rhs_eval"],
      "sde", c["vdp_sde_n4194304"].get("em_ms_per_step"), "leap", c["vdp_sde_n4194304"].get("tauleap_ms_steady"), [round(s["ms"], 2) for s in c["vdp_sde_n4194304"]["tauleap_sweeps"]])
P
