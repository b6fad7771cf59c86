#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py -x -q -m gpu -k "native or side_arrays or reference_trace or graphed or invariants" 2>&1 | tail -3
tools/gpu_exp_build.sh default "-DBGYM_GATHER_CTAS=16" "-DBGYM_GATHER_WARPS=2 -DBGYM_GATHER_CTAS=8"
BGYM_NVCC_EXTRA="-DBGYM_TILE_CLOCK" python -c "import balatro_gym_b200 as b; b.build(force=True)"; timeout 300 python bench.py --steps 130 --warmup 3 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 2>&1 >/dev/null | grep "bgym t" | sed -n 2,3p
