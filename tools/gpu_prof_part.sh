#!/bin/bash
# ncu --set full capture of the four launches of one env-step (main pass + three gather passes) at bench size
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"env_step_(main|list)" -s 400 -c 10 -f -o gpurun_out/prof_part python bench.py --steps 20 --warmup 3 --burn-in 100 --no-cpu-baseline --no-hands --no-ppo --no-facade --e2e-steps 3 > gpurun_out/ncu_part.log 2>&1
tail -3 gpurun_out/ncu_part.log
