"""Env-step replayed from a CUDA graph vs launched call by call (2^20 envs)."""
import torch
from balatro_gym_b200 import BalatroVecEnv
n = 1 << 20
env = BalatroVecEnv(n, seed=1); env.reset(); env.randomize_c3(1)
for _ in range(150): env.step(random_policy=True)
def timed(fn, K=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K
def eager():
    env.sample_actions(seed=7); env.step(env.actions, want_info=False)
print("eager  sampler+step: %.4f ms -> %.3e env-steps/s" % (timed(eager), n / timed(eager) * 1e3))
g = env.graphed_rollout_step("sampler", seed=7)
t = timed(g); print("graph  sampler+step: %.4f ms -> %.3e env-steps/s" % (t, n / t * 1e3))
t = timed(lambda: env.step(random_policy=True, want_info=False)); print("eager  fused: %.4f ms -> %.3e" % (t, n / t * 1e3))
gf = env.graphed_rollout_step("fused")
t = timed(gf); print("graph  fused: %.4f ms -> %.3e" % (t, n / t * 1e3))
print("episodes so far", int(env.state_field("episode").long().sum()), "terminated now", int(env.terminated.sum()))
