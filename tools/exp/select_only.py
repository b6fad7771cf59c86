"""Experiment: step-kernel time when every env takes a SELECT toggle (converged, tiny code path)
versus the random-legal mix.  Lower bound for a category-partitioned design."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import balatro_gym_b200 as b
n = 1 << 20
env = b.BalatroVecEnv(n, seed=1, autoreset=True)
env.reset(); env.randomize_c3(seed=1)
for _ in range(150):
    env.sample_actions(seed=1); env.step(env.actions, want_info=False)
def timeit(fn, k=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
sel = torch.full((n,), 3, dtype=torch.int32, device="cuda")
t_sel = timeit(lambda: env.step(sel, want_info=False))
def mixed():
    env.sample_actions(seed=1); env.step(env.actions, want_info=False)
t_mix = timeit(mixed)
t_samp = timeit(lambda: env.sample_actions(seed=1))
print(f"variant {os.environ.get('BGYM_VARIANT','default')}: select-only step {t_sel:.3f} ms, sampler {t_samp:.3f} ms, sampler+mixed step {t_mix:.3f} ms")
