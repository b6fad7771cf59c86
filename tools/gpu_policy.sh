#!/bin/bash
# fused tcgen05 policy forward: correctness test (under a short timeout: a wrong barrier phase would hang), then the PPO bench block
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "fused_policy" 2>&1 | grep -v "^E    \+" | tail -15
