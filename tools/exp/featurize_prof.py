import torch
from balatro_gym_b200 import BalatroVecEnv
from balatro_gym_b200.rollout import featurize
v = BalatroVecEnv(1 << 19, seed=1); v.reset()
for _ in range(30): v.step(random_policy=True)
out = torch.empty((1 << 19, 448), dtype=torch.bfloat16, device="cuda")
for _ in range(5): featurize(v.obs_buf, out=out)
torch.cuda.synchronize()
