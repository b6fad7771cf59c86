#!/bin/bash
# Round-2 record run: GPU suite, smoke, both bench arms, ncu captures of the non-step kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=json.load(open('gpurun_out/bench_reference.json')); print('value %.3e frac %.3f kernel_ms %.3f fused %.3e e2e %.3e hands %.3e ppo %.3e (policy %.3f ms, library %.3f ms) | reference arm %.3e on %s cores' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['e2e']['value'], d['hands']['value'], d['ppo_rollout']['value'], d['ppo_rollout']['breakdown_ms']['policy_forward_fused_tcgen05'], d['ppo_rollout']['policy_forward_library_gemms_ms'], r['value'], r['cpu_baseline']['cores']))" || tail -5 gpurun_out/bench.err
# tools/gpu_prof_small.sh   (run separately)
ls gpurun_out | wc -l
