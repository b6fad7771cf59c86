#!/bin/bash
# ncu --set full of the non-step kernels, one launch each: joker interpreter, hands5, the fused tcgen05 policy forward, masked sample, sampler
mkdir -p gpurun_out
for k in score_hands_kernel policy_mlp_kernel masked_sample_kernel sample_actions_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 3 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 5 --warmup 3 --burn-in 5 --no-cpu-baseline --no-facade --e2e-steps 3 --ppo-steps 2 > gpurun_out/ncu_small.log 2>&1
  tail -1 gpurun_out/ncu_small.log | cut -c1-200
done
