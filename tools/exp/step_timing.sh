# BGYM_STEP_TIMING=1: phases of a step (synchronising); =2: every kernel of a step launched alone.  Lines 5-6 of each run are from
# the sampler + step loop (burn-in and timed steps), the last lines from the fused-policy loop
for t in 1 2; do BGYM_STEP_TIMING=$t timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 2>&1 >/dev/null | grep "bgym timing" | sed -n '4,5p;$p'; done
