#!/bin/bash
# multi-GPU measurements: bash scripts/diag_ngpu.sh <N> <tag>   (bench at N GPUs, sync-cost microbench, parity test)
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python scripts/sync_cost.py > gpurun_out/${TAG}_sync1.log 2>&1; cat gpurun_out/${TAG}_sync1.log | grep world
timeout 300 $TR --master-port 29511 scripts/sync_cost.py > gpurun_out/${TAG}_sync$N.log 2>&1; grep world gpurun_out/${TAG}_sync$N.log
run() { name=$1; shift; timeout 400 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().split("\n")[-1])
    print("   ms/step %.4f sustained %.4f per-rank %s variant %s" % (d["ms_per_step"], d["ms_per_step_sustained"], ["%.4f" % x for x in d["per_rank_ms_per_step"]], d["run"]["step_kernel_variant"]))
    c = d.get("configs") or {}
    for k, v in c.items():
        if isinstance(v, dict): print("   ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("ms_per_rhs_eval", "rk4_ms_per_step", "em_ms_per_step", "tauleap_ms_steady")}, v.get("rkf45", {}).get("ms_per_accepted_step"), round(v.get("roofline", {}).get("frac", 0), 3))
        else: print("   ", k, v)
except Exception as e:
    print("   no line:", e)
PY
}
run noconf --no-configs
run full
run nopeers --no-configs --no-peers
