#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
VARIANTS="8" bash tools/gpu_sweep.sh
BGYM_SERIAL_GATHER=1 BGYM_VARIANT=8 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('V8-serial value %.3e frac %.3f kernel_ms %.3f fused %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value']))"
BGYM_SERIAL_GATHER=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 24 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --burn-in 100 --no-cpu-baseline --no-hands --e2e-steps 3 > gpurun_out/ncu_bench.log 2>&1
grep -E "env_step|sample_actions" gpurun_out/launches.csv | awk -F'","' '{print $5, $NF}' | sort | head -24
