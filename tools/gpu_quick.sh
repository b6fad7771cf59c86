#!/bin/bash
# Quick GPU visit: parity tests + bench (variant 0/1) + ncu full capture of the step kernel.
mkdir -p gpurun_out
ls -la oracle/_ref oracle/_ref/balatro_gym 2>&1 | head -8 > gpurun_out/ref_ls.txt
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
BGYM_VARIANT=0 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v0.json')); print('V0 value %.3e frac %.3f kernel_ms %.3f fused %.3e e2e %.3e hands %.3e hfrac %.3f cpu %s' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['e2e']['value'], d['hands']['value'], d['hands']['roofline']['frac'], d['cpu_baseline']))"; tail -3 gpurun_out/bench_v0.err
BGYM_VARIANT=1 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v1.json')); print('V1 value %.3e frac %.3f kernel_ms %.3f fused %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value']))"; tail -3 gpurun_out/bench_v1.err
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 40 -c 1 -f -o gpurun_out/prof_step python bench.py --steps 20 --warmup 3 --burn-in 60 --no-cpu-baseline --no-hands --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_hands -s 2 -c 1 -f -o gpurun_out/prof_hands python bench.py --steps 5 --warmup 3 --burn-in 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full_hands.log 2>&1
