"""Where does the e2e step's time go?  Times the pieces of HostMirror.step on their own (B200 box)."""
import time
import torch
import balatro_gym_b200 as b
from balatro_gym_b200 import layout as L

n = 1 << 20
dev = torch.device("cuda:0")
env = b.BalatroVecEnv(n, device=dev, seed=1, generator="c4")
env.reset()
for _ in range(120):
    env.sample_actions(seed=3); env.step(env.actions, want_info=False)
m = b.HostMirror(env)
m.pull_all()
lib = env.lib
st = torch.cuda.current_stream(dev)

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

snap = m._snap[0]
env.sample_actions(seed=3); env.step(env.actions, want_info=False)
keep = env.obs_dirty.clone()
def pack():
    env.obs_dirty.copy_(keep)
    lib.bgym_pack_dirty_obs(env._obs.data_ptr(), env.obs_dirty.data_ptr(), snap["staging"].data_ptr(), m._scratch.data_ptr(), n, n, st.cuda_stream)
t_pack = timed(pack)
cnt, flagged = m.delta_counts(0)
print("dirty records %d (%.3f), with shop chunks %d (%.3f)  pack %.3f ms" % (cnt, cnt / n, flagged, flagged / n, t_pack))
def scat_host(): lib.bgym_scatter_dirty_obs(snap["staging"].data_ptr(), n, m.core.data_ptr(), m.shop.data_ptr(), st.cuda_stream)
t = timed(scat_host); print("scatter -> pinned host  %.3f ms  %.1f GB/s" % (t, (cnt * 128 + flagged * 32) / t / 1e6))
dcore, dshop = torch.empty((n, 128), dtype=torch.uint8, device=dev), torch.empty((n, 32), dtype=torch.uint8, device=dev)
def scat_dev(): lib.bgym_scatter_dirty_obs(snap["staging"].data_ptr(), n, dcore.data_ptr(), dshop.data_ptr(), st.cuda_stream)
t = timed(scat_dev); print("scatter -> device       %.3f ms" % t)
hs = torch.empty(16 + n * 4 + cnt * 176, dtype=torch.uint8, pin_memory=True)
def dma_staged(): hs.copy_(snap["staging"][:hs.numel()], non_blocking=True)
t = timed(dma_staged); print("DMA of the staged block (idx for all + %d records = %.1f MB)  %.3f ms  %.1f GB/s" % (cnt, hs.numel() / 1e6, t, hs.numel() / t / 1e6))
def dma_dense():
    m.sel.copy_(env.sel, non_blocking=True); m.reward.copy_(env.reward, non_blocking=True); m.terminated.copy_(env.terminated, non_blocking=True)
t = timed(dma_dense); print("DMA sel+reward+term (%.1f MB)  %.3f ms  %.1f GB/s" % (25 * n / 1e6, t, 25 * n / t / 1e6))
def h2d(): m._d_act.copy_(m.actions, non_blocking=True)
t = timed(h2d); print("H2D actions  %.3f ms" % t)
def d2h_act(): m.actions.copy_(env.actions, non_blocking=True)
t = timed(d2h_act); print("D2H actions  %.3f ms" % t)
def step(): env.sample_actions(seed=3); env.step(env.actions, want_info=False)
t = timed(step, 50); print("sampler + step  %.3f ms" % t)
# the mirror loop as bench.py runs it
def loop(k):
    for t_ in range(k):
        env.sample_actions(seed=7); m.actions.copy_(env.actions, non_blocking=True); st.synchronize(); m.step()
    m.wait()
loop(3)
t0 = time.perf_counter(); loop(20); dt = (time.perf_counter() - t0) / 20
print("mirror loop %.3f ms/step  %.3e env-steps/s" % (dt * 1e3, n / dt))
# without the device->host->device action hand-off (actions stay where the device policy wrote them)
def loop2(k):
    for t_ in range(k):
        env.sample_actions(seed=7); m.actions.copy_(env.actions, non_blocking=True); m.step()
    m.wait()
loop2(3)
t0 = time.perf_counter(); loop2(20); dt = (time.perf_counter() - t0) / 20
print("mirror loop, no host sync on the action path %.3f ms/step  %.3e env-steps/s" % (dt * 1e3, n / dt))
