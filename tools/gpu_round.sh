#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both staging variants), ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
BGYM_VARIANT=0 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.err; tail -c 3000 gpurun_out/bench_v0.json; tail -3 gpurun_out/bench_v0.err
BGYM_VARIANT=1 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; tail -c 1500 gpurun_out/bench_v1.json; tail -3 gpurun_out/bench_v1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --burn-in 30 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 40 -c 2 -f -o gpurun_out/prof_step python bench.py --steps 20 --warmup 3 --burn-in 60 --no-cpu-baseline --no-hands --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_hands -s 2 -c 1 -f -o gpurun_out/prof_hands python bench.py --steps 5 --warmup 3 --burn-in 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full_hands.log 2>&1
ls -la gpurun_out | head -30
