#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | grep -v "^E    \+" | tail -8
timeout 300 python - <<'PY'
import torch, bench, json, argparse, time
from balatro_gym_b200 import dist as bdist
from balatro_gym_b200.rollout import make_policy, pack_policy_weights, policy_forward_fused
rank, lr, ws = bdist.init_process_group("nccl")
dev = torch.device("cuda:0")
n = 1 << 19
a = (torch.rand((n, 448), device=dev) * 2).to(torch.bfloat16)
pol = make_policy(device=dev, seed=0)
w, b = pack_policy_weights(pol.state_dict(), dev)
lg = torch.empty((n, 60), device=dev); vl = torch.empty(n, device=dev)
for _ in range(3): policy_forward_fused(a, w, b, lg, vl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): policy_forward_fused(a, w, b, lg, vl)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2 * n * (256*128 + 128*64 + 64*32 + 256*512 + 512*512 + 2*(512*256 + 256*256) + 256*64 + 256*16)
print("fused MLP forward at 2^19 envs: %.3f ms  %.1f TFLOP/s (padded shapes)" % (ms, fl / ms / 1e9))
args = argparse.Namespace(ppo_envs=1 << 19, ppo_steps=16)
r = bench.bench_ppo_rollout(torch, bdist, dev, args, 0, 1)
print("ppo %.3e env-steps/s  %.3f ms/step  %s" % (r["value"], r["ms_per_step"], json.dumps({k: round(v, 3) for k, v in r["breakdown_ms"].items()})))
PY
