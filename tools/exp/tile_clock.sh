# level-1 tile timeline of the production loop (-DBGYM_TILE_CLOCK diagnostic build): per list, when its first tile started and
# its last tile ended (us from the level's first tile), mean / max tile latency; two consecutive prints
BGYM_NVCC_EXTRA="-DBGYM_TILE_CLOCK" python -c "import balatro_gym_b200 as b; b.build(force=True)"
timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 2>&1 >/dev/null | grep "bgym tiles" | sed -n '4,5p'
