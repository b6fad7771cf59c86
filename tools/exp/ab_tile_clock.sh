# tile timelines (-DBGYM_TILE_CLOCK) of the last commit's kernels (tools/exp/_ab_old) and the working tree's, same box
mkdir -p /tmp/ab; cp -r balatro_gym_b200/csrc /tmp/ab/csrc_new
run() { BGYM_NVCC_EXTRA="-DBGYM_TILE_CLOCK" python -c "import balatro_gym_b200 as b; b.build(force=True)"; timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 2>&1 >/dev/null | grep "bgym tiles" | sed -n '4,5p' | sed "s/^/[$1] /"; }
cp tools/exp/_ab_old/* balatro_gym_b200/csrc/; run old
cp /tmp/ab/csrc_new/* balatro_gym_b200/csrc/; run new
