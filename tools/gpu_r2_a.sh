#!/bin/bash
# round-2 visit A: generator parity on the GPU + a short bench with the device-side generator
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "native_rollout or generator or golden or replays" > gpurun_out/pytest_gen.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gen.log
tail -5 gpurun_out/pytest_gen.log
BGYM_STEP_TIMING=1 python bench.py --steps 128 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err; grep "bgym timing" gpurun_out/bench_timing.err | tail -3
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-ppo --no-facade > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; python -c "
import json; d=json.load(open('gpurun_out/bench_a.json')); print('value %.3e frac %.3f kernel_ms %.3f fused %.3e e2e %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['e2e']['value'])); print(json.dumps(d['timed_window']))" || tail -5 gpurun_out/bench_a.err
