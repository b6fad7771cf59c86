#!/bin/bash
# One full GPU-box visit for the round: parity tests, smoke, bench (our arm + reference arm),
# ncu launch list of the bench command, ncu --set full of the step kernels and the hands kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=json.load(open('gpurun_out/bench_reference.json')); print('value %.3e frac %.3f kernel_ms %.3f fused %.3e e2e %.3e hands %.3e hfrac %.3f | reference arm %.3e on %s cores' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['e2e']['value'], d['hands']['value'], d['hands']['roofline']['frac'], r['value'], r['cpu_baseline']['cores']))" || tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --burn-in 30 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"env_step_(main|gather)" -s 400 -c 4 -f -o gpurun_out/prof_step python bench.py --steps 20 --warmup 3 --burn-in 100 --no-cpu-baseline --no-hands --e2e-steps 3 > gpurun_out/ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_hands -s 2 -c 1 -f -o gpurun_out/prof_hands python bench.py --steps 5 --warmup 3 --burn-in 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_hands.log 2>&1
ls gpurun_out | head -30
