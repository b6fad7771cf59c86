#!/bin/bash
# build variants (nvcc -D flags) compared on the production step loop: value, kernel_ms (CUDA events around the step's launches)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "default" ]; then BGYM_NVCC_EXTRA="" python -c "import balatro_gym_b200 as b; b.build(force=True)" || exit 1
  else BGYM_NVCC_EXTRA="$v" python -c "import balatro_gym_b200 as b; b.build(force=True)" || { echo "[$v] build failed"; continue; }; fi
  timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print('[$v] value %.3e kernel_ms %.4f fused %.3e graph %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['graph_replay']['sampler_plus_step']))" || tail -3 gpurun_out/bench_v.err
done
