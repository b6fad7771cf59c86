#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for v in 8 0; do BGYM_VARIANT=$v python tools/exp_select_only.py; done
VARIANTS="8" bash tools/gpu_sweep.sh
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --burn-in 100 --no-cpu-baseline --no-hands --e2e-steps 3 > gpurun_out/ncu_bench.log 2>&1
grep -E "env_step|sample_actions" gpurun_out/launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -20
