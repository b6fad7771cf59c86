#!/bin/bash
# round-2 visit B: parity of the list-partitioned step + timing split + short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "native_rollout or generator or replays or ragged or host_buffer or without_observations or full_size_rollout or refresh_observations" > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b.log
tail -15 gpurun_out/pytest_b.log
BGYM_STEP_TIMING=1 timeout 300 python bench.py --steps 128 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err; grep "bgym timing" gpurun_out/bench_timing.err | tail -3
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-ppo --no-facade --no-hands > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print('value %.3e frac %.3f kernel_ms %.3f fused %.3e graph %.3e e2e %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['graph_replay']['sampler_plus_step'], d['e2e']['value']))" || tail -5 gpurun_out/bench_b.err
