#!/bin/bash
# ncu --set full capture of the launches of ONE env-step (main pass + seven list kernels + the level-2 kernel) at bench size, plus the
# five-card hands kernel; read here with tools/ncu_summary.py (--traffic-json writes profiles/r02_traffic.json)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"env_step_(main|list|level)" -s 396 -c 9 -f -o gpurun_out/prof_part python bench.py --steps 20 --warmup 3 --burn-in 100 --no-cpu-baseline --no-hands --no-ppo --no-facade --e2e-steps 3 > gpurun_out/ncu_part.log 2>&1
tail -1 gpurun_out/ncu_part.log
ncu --set full --clock-control none -k regex:"^score_hands5_kernel" -s 3 -c 1 -f -o gpurun_out/prof_score_hands5_kernel python bench.py --steps 5 --warmup 3 --burn-in 5 --no-cpu-baseline --no-facade --no-ppo --e2e-steps 3 > gpurun_out/ncu_small.log 2>&1
tail -1 gpurun_out/ncu_small.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --burn-in 30 --no-cpu-baseline --e2e-steps 3 --no-facade > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log
