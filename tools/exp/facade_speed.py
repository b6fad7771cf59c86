"""Steps/s of the N=1 Gymnasium facade and of small SB3-style slabs (host-facing paths)."""
import time, numpy as np
from balatro_gym_b200.env import BalatroEnv
from balatro_gym_b200.sb3_vec_env import BalatroSB3VecEnv
env = BalatroEnv(seed=3)
obs, _ = env.reset(seed=3)
rng = np.random.default_rng(0)
def run(nsteps):
    global obs
    t0 = time.perf_counter()
    for _ in range(nsteps):
        a = int(rng.choice(np.flatnonzero(obs["action_mask"])))
        obs, r, term, trunc, info = env.step(a)
        if term:
            obs, _ = env.reset()
    return nsteps / (time.perf_counter() - t0)
run(200)
print("BalatroEnv facade (N=1): %.0f steps/s" % run(3000))
for n in (8, 64, 1024):
    v = BalatroSB3VecEnv(n, seed=1)
    o = v.reset()
    def vrun(k):
        global o
        t0 = time.perf_counter()
        for _ in range(k):
            m = o["action_mask"]
            # uniform legal action per env (vectorised on the host)
            u = rng.random(m.shape) * m
            o, r, d, i = v.step(u.argmax(axis=1))
        return k * n / (time.perf_counter() - t0)
    vrun(20)
    print("BalatroSB3VecEnv n=%d: %.0f env-steps/s" % (n, vrun(300 if n < 1024 else 100)))
