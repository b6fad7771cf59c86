# per-kernel step timings (BGYM_STEP_TIMING=2, every kernel alone) and phase timings (=1) of build variants
for v in "$@"; do
  if [ "$v" = "default" ]; then f=""; else f="$v"; fi
  BGYM_NVCC_EXTRA="$f" python -c "import balatro_gym_b200 as b; b.build(force=True)" || { echo "[$v] build failed"; continue; }
  for t in 2 1; do echo -n "[$v] "; BGYM_STEP_TIMING=$t timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 2>&1 >/dev/null | grep "bgym timing" | sed -n '5p'; done
done
