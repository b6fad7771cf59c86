#!/bin/bash
# fused tcgen05 policy forward: rollout tests (under a timeout: a wrong barrier phase would hang), kernel timing, PPO bench block
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | grep -v "^E    \+" | tail -6
tools/gpu_policy_time.sh 2>&1 | tail -2
timeout 300 python - <<'PY'
import torch, bench, json, argparse
from balatro_gym_b200 import dist as bdist
rank, lr, ws = bdist.init_process_group("nccl")
dev = torch.device("cuda:0")
args = argparse.Namespace(ppo_envs=1 << 19, ppo_steps=16)
r = bench.bench_ppo_rollout(torch, bdist, dev, args, 0, 1)
print("ppo %.3e env-steps/s  %.3f ms/step  %s" % (r["value"], r["ms_per_step"], json.dumps({k: round(v, 3) for k, v in r["breakdown_ms"].items()})))
PY
