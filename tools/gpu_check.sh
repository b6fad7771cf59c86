#!/bin/bash
# Parity check of a kernel change + A/B timing: golden-trace replay and native CUDA-vs-oracle tests, a short soak on both
# launch structures, then the production step loop.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python tools/soak_parity.py --envs 16384 --steps 300 2>&1 | tail -2
python tools/soak_parity.py --envs 4096 --steps 300 --one-launch 2>&1 | tail -2
bash tools/gpu_exp_build.sh "$@"
