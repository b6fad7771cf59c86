#!/bin/bash
cat > /tmp/pt.py <<'PY'
import torch, os
from balatro_gym_b200.rollout import make_policy, pack_policy_weights, policy_forward_fused
dev = torch.device("cuda:0")
n = 1 << 19
import balatro_gym_b200 as bb
env = bb.BalatroVecEnv(n, device=dev, seed=1, generator="c4"); env.reset()
for _ in range(60): env.step(random_policy=True, want_info=False)
a = env.obs_buf
pol = make_policy(device=dev, seed=0)
w, b = pack_policy_weights(pol.state_dict(), dev)
lg = torch.empty((n, 60), device=dev); vl = torch.empty(n, device=dev)
for _ in range(3): policy_forward_fused(a, w, b, lg, vl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): policy_forward_fused(a, w, b, lg, vl)
e1.record(); torch.cuda.synchronize()
print("fused MLP forward at 2^19 envs: %.3f ms" % (e0.elapsed_time(e1) / 10))
PY
# cta_group::2 variant of the policy kernel: correctness under a timeout, then timing (both variants)
BGYM_POLICY_CTAS=2 timeout 120 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k fused 2>&1 | grep -v "^E    \+" | tail -6
echo "--- timing, pairs"
BGYM_POLICY_CTAS=2 PYTHONPATH=. timeout 120 python /tmp/pt.py
BGYM_POLICY_CTAS=2 PYTHONPATH=. BGYM_POLICY_CLOCK=1 timeout 120 python /tmp/pt.py 2>&1 | grep clocks | tail -1
echo "--- timing, single CTAs"
PYTHONPATH=. timeout 120 python /tmp/pt.py
