TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 300 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 20 --warmup 5 --no-configs --no-cpu "$@" > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b_$name.json").read().strip().split("\n")[-1])
    print("   ms/step %.4f sustained %.4f per-rank %s variant %s clocks %s" % (d["ms_per_step"], d["ms_per_step_sustained"], d["per_rank_ms_per_step"], d["run"]["step_kernel_variant"], d["clocks"]["per_gpu_sm_mhz"]))
except Exception as e:
    print("   no line:", e)
PY
}
run default
run default2
run skew30 --skew-ms 30
run lean0 --opt lean=0
run tma0 --opt tma=0
run nopeers --no-peers
run steps200 --steps 200
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2b_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -30 gpurun_out/r2b_pytest_multi.log | cut -c1-220
