#!/bin/bash
# the round's single-GPU record in one box call (development tool): GPU tests, smoke, the bench as the driver runs it, its launch list
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2v_pytest_gpu.txt; cat gpurun_out/r2v_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench20.err | grep "^{" > gpurun_out/r2v_bench_n1_steps20.json
python bench.py 2>gpurun_out/bench.err | grep "^{" > gpurun_out/r2v_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2v_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
python - <<'P'
import json
for f in ("gpurun_out/r2v_bench_n1_steps20.json", "gpurun_out/r2v_bench_n1.json"):
    d = json.loads(open(f).read()); c = d["configs"]
    print(f, d["ms_per_step"], d.get("ms_per_step_sustained"), d["roofline"]["frac"], d["roofline"].get("frac_sustained"), d["e2e"]["value"], d["clocks"]["sm_mhz"], d["clocks"]["samples_in_timed_region"], d["clocks"]["reasons"])
    nb = c["nbody_rkf45_n131072"]; print(" nbody", nb["ms_per_rhs_eval"], nb["ms_per_rhs_eval_ordered_kernel"], nb["rkf45"], nb["roofline"]["frac"])
    print(" lowrank", c["lowrank64_n1048576"]["ms_per_rhs_eval"], c["lowrank64_n1048576"]["roofline"]["frac"], " sde", c["vdp_sde_n4194304"]["em_ms_per_step"], c["vdp_sde_n4194304"]["roofline"]["frac"], c["vdp_sde_n4194304"]["tauleap_ms_steady"])
    print(" cpu", d["cpu_baseline"]["value"], " parity", d["parity"])
P
