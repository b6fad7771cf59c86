#!/bin/bash
# Toggle / selection side arrays: full GPU suite, short bench, phase split of the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print('value %.3e frac %.3f kernel_ms %.3f fused %.3e e2e %.3e pcie %.1f GB/s dirty %.3f graph %s' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['e2e']['value'], d['e2e']['pcie_gbs_rank0'], d['e2e']['rewritten_record_frac'], d['graph_replay']))" || tail -5 gpurun_out/bench_q.err
BGYM_STEP_TIMING=1 timeout 300 python bench.py --steps 130 --warmup 3 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 2>&1 >/dev/null | grep "bgym timing" | tail -2
BGYM_STEP_TIMING=2 timeout 300 python bench.py --steps 130 --warmup 3 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 2>&1 >/dev/null | grep "bgym timing" | tail -2
