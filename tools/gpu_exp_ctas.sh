#!/bin/bash
# per-kernel (serial) and per-phase (concurrent) timing of the step's launches in the non-fused bench loop
mkdir -p gpurun_out
for v in "$@"; do
  if [ -n "$v" ] && [ "$v" != "default" ]; then BGYM_NVCC_EXTRA="$v" python -c "import balatro_gym_b200 as b; b.build(force=True)" || exit 1; fi
  for t in 2 1; do
  BGYM_STEP_TIMING=$t timeout 300 python bench.py --steps 128 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade > /dev/null 2> gpurun_out/t$t.err
  echo "[$v] $(grep 'bgym timing' gpurun_out/t$t.err | sed -n 3p | cut -c1-260)"
  done
done
