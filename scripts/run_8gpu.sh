#!/bin/bash
# the round's 8-GPU measurements in one box call (development tool): quick parity check, bench at 8 and 4 ranks, exchange cost
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
MGPU_QUICK=1 $T --nproc-per-node 8 --master-port 29721 scripts/mgpu_check.py 2>&1 | grep -v "^W\|^\[W\|^\*\*\*\|OMP_NUM" > gpurun_out/r2u_mgpu_check_8gpu.log
tail -3 gpurun_out/r2u_mgpu_check_8gpu.log
$T --nproc-per-node 8 --master-port 29722 bench.py --gpus 8 --steps 20 --warmup 5 2>&1 | grep "^{" > gpurun_out/r2u_bench_n8.json
$T --nproc-per-node 4 --master-port 29723 bench.py --gpus 4 --steps 20 --warmup 5 2>&1 | grep "^{" > gpurun_out/r2u_bench_n4.json
$T --nproc-per-node 8 --master-port 29724 scripts/sync_cost.py 2>&1 | grep "^world" > gpurun_out/r2u_sync_cost_8gpu.txt
cat gpurun_out/r2u_sync_cost_8gpu.txt
python - <<'P'
import json
for f in ("gpurun_out/r2u_bench_n8.json", "gpurun_out/r2u_bench_n4.json"):
    try:
        d = json.loads(open(f).read())
        c = d["configs"]
        print(d["n_gpus"], "ms/step", d["ms_per_step"], "sustained", d.get("ms_per_step_sustained"), "nbody ms/eval", c["nbody_rkf45_n131072"]["ms_per_rhs_eval"],
              "rkf45 ms/step", c["nbody_rkf45_n131072"]["rkf45"]["ms_per_accepted_step"], "lowrank", c["lowrank64_n1048576"]["ms_per_rhs_eval"],
              "sde", c["vdp_sde_n4194304"].get("em_ms_per_step"), "leap", c["vdp_sde_n4194304"].get("tauleap_ms_best"))
    except Exception as e:
        print(f, "unreadable:", e)
P
