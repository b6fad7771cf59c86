"""GPU tests of the rows SURVEY §8(f) marks "next": the SB3 VecEnv adapter against a DummyVecEnv-style
loop over unmodified reference envs, the determinism / action-masking validators
(balatro_env_2.py:1737-1796) run against the device env, and checkpoint round trips."""
import numpy as np
import pytest

from balatro_gym_b200 import layout as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def test_sb3_vec_env_matches_dummyvecenv_over_reference(torch, reference):
    """env i of BalatroSB3VecEnv(shuffle='reference') == `BalatroEnv(seed=seed+i)` of the reference
    driven with DummyVecEnv's auto-reset protocol: identical observations (incl. the first one after
    each auto-reset: the MT19937 deck stream continues across episodes), rewards, dones,
    terminal_observation and Monitor's episode summary.  An env stops being compared once it reaches
    the shop (shop draws come from other MT streams that only replay mode carries)."""
    from balatro_gym_b200.sb3_vec_env import BalatroSB3VecEnv
    n, seed = 12, 101
    venv = BalatroSB3VecEnv(n, seed=seed, shuffle="reference")
    refs = [reference.BalatroEnv(seed=seed + i) for i in range(n)]
    robs = [e.reset()[0] for e in refs]           # DummyVecEnv.reset(): env.reset(seed=None)
    obs = venv.reset()
    live = np.ones(n, dtype=bool)
    rng = np.random.default_rng(5)
    ep_ret = np.zeros(n); ep_len = np.zeros(n, dtype=int)
    episodes_checked = 0
    for t in range(600):
        for i in np.flatnonzero(live):
            for k in L.OBS_KEYS:
                assert np.array_equal(np.asarray(robs[i][k]), obs[k][i]), (t, i, k)
        masks = venv.action_masks()
        actions = np.zeros(n, dtype=np.int64)
        for i in range(n):
            legal = np.flatnonzero(robs[i]["action_mask"]) if live[i] else np.flatnonzero(masks[i])
            a = int(rng.choice(legal))
            if rng.random() < 0.03:
                a = int(rng.integers(0, 60))       # occasionally an arbitrary (often illegal) action
            if a == 47 and live[i]:
                a = 45                             # the boss draw comes from the module-global RNG (replay mode only)
            actions[i] = a
        obs, rew, dones, infos = venv.step(actions)
        assert rew.dtype == np.float32 and dones.dtype == bool and len(infos) == n
        for i in np.flatnonzero(live):
            ro, rr, rt, rtr, ri = refs[i].step(int(actions[i]))
            if refs[i].state.phase == 1:           # entered the shop: stop comparing this env
                live[i] = False
                continue
            ep_ret[i] += rr; ep_len[i] += 1
            assert np.float32(rr) == rew[i], (t, i, rr, rew[i])
            assert bool(rt or rtr) == bool(dones[i]), (t, i)
            assert ("error" in ri) == ("error" in infos[i]), (t, i, ri, infos[i])
            if rt:
                term = infos[i]["terminal_observation"]
                for k in L.OBS_KEYS:
                    assert np.array_equal(np.asarray(ro[k]), term[k]), (t, i, k)
                assert infos[i]["episode"]["l"] == ep_len[i]
                assert abs(infos[i]["episode"]["r"] - ep_ret[i]) < 1e-5
                assert infos[i]["TimeLimit.truncated"] is False
                ep_ret[i] = 0; ep_len[i] = 0
                episodes_checked += 1
                ro = refs[i].reset()[0]            # DummyVecEnv.step_wait(): reset, return the new obs
            robs[i] = ro
        if not live.any():
            break
    assert episodes_checked >= 5, episodes_checked


def test_sb3_vec_env_guards_and_protocol(torch):
    """SafeBalatroEnv's guards (train_balatro_fixed.py:240-258) and the VecEnv helper surface."""
    from balatro_gym_b200.sb3_vec_env import BalatroSB3VecEnv, next_episode_seed
    n = 64
    venv = BalatroSB3VecEnv(n, seed=7, max_invalid_actions=5, max_episode_steps=40)
    obs = venv.reset()
    assert set(obs.keys()) == set(L.OBS_KEYS) and obs["hand"].shape == (n, 8) and obs["action_mask"].shape == (n, 60)
    assert obs["phase"].tolist() == [2] * n
    # action 0 (PLAY_HAND) is illegal in BLIND_SELECT: rejected with -1.0, fifth rejection ends the episode with -50
    for k in range(5):
        obs, rew, dones, infos = venv.step(np.zeros(n, dtype=np.int64))
        if k < 4:
            assert (rew == -1.0).all() and not dones.any() and all("error" in i for i in infos)
    assert (rew == -50.0).all() and dones.all()
    assert all(i["invalid_action_termination"] and "terminal_observation" in i and i["episode"]["l"] == 5 for i in infos)
    assert np.allclose([i["episode"]["r"] for i in infos], -54.0)
    # the seed chain is the kernels' own
    assert venv.get_attr("rng_seed")[3] == next_episode_seed(np.array([7 + 3], dtype=np.uint32))[0]
    # time limit: legal random play for 40 steps -> truncated envs carry TimeLimit.truncated
    rng = np.random.default_rng(0)
    saw_trunc = False
    for t in range(45):
        masks = np.stack(venv.env_method("action_masks"))
        acts = np.array([rng.choice(np.flatnonzero(m)) for m in masks])
        obs, rew, dones, infos = venv.step(acts)
        for i in np.flatnonzero(dones):
            if infos[i].get("max_steps_reached"):
                assert infos[i]["TimeLimit.truncated"] == (not infos[i].get("failed", False))
                saw_trunc = True
    assert saw_trunc
    assert venv.env_is_wrapped(object) == [False] * n and venv.get_images() == [None] * n
    venv.set_attr("money", 77, indices=[0, 1])
    assert venv.get_attr("money", indices=[0, 1, 2])[:2] == [77, 77]
    venv.close()


def test_validators_on_the_facade(torch):
    """The reference's own behavioural checks (BalatroEnvValidator, balatro_env_2.py:1737-1796)."""
    from balatro_gym_b200.env import BalatroEnv
    from balatro_gym_b200.validate import BalatroEnvValidator
    assert BalatroEnvValidator.validate_determinism(BalatroEnv, seed=42, steps=100)
    assert BalatroEnvValidator.validate_action_masking(BalatroEnv(seed=42))


def test_validators_batched(torch):
    from balatro_gym_b200 import validate
    assert validate.validate_determinism_vec(num_envs=8192, seed=42, steps=120)
    assert validate.validate_action_masking_vec(num_envs=8192, seed=42)
    assert validate.validate_checkpoint_roundtrip(num_envs=8192, seed=3, steps=60)


def test_graphed_rollout_step_matches_eager(torch):
    """The env-step replayed from a CUDA graph (sampler or fused policy) leaves exactly the state, observations,
    rewards and actions that call-by-call launches leave."""
    from balatro_gym_b200 import BalatroVecEnv
    n = 1 << 17            # above the small-slab threshold: the multi-pass step with its forked gather streams
    for policy in ("fused", "sampler"):
        a, b = BalatroVecEnv(n, seed=9), BalatroVecEnv(n, seed=9)
        for v in (a, b):
            v.reset()
            v.randomize_c3(2)

        def eager():
            if policy == "fused":
                a.step(random_policy=True, want_info=False)
            else:
                a.step(a.sample_actions(seed=5), want_info=False)   # step numbers 0, 1, 2, ... like the device counter
        replay = b.graphed_rollout_step(policy, seed=5)     # runs two real warm-up steps; the capture pass runs nothing
        eager(); eager()
        for _ in range(40):
            replay()
            eager()
        torch.cuda.synchronize()
        if policy == "sampler":
            assert int(b._step_ctr.item()) == 42
        for name in ("hot", "cold", "obs_buf", "reward", "terminated", "actions"):
            assert torch.equal(getattr(a, name), getattr(b, name)), (policy, name)


def test_plain_c_caller_runs(torch, tmp_path):
    """examples/host_loop.c against the library on the GPU: a C program steps 4096 envs from host buffers."""
    import os
    import subprocess
    from conftest import REPO
    from balatro_gym_b200 import _lib
    exe = str(tmp_path / "host_loop")
    subprocess.check_call(["gcc", "-Wall", "-I", os.path.join(REPO, "include"), "-o", exe,
                           os.path.join(REPO, "examples", "host_loop.c"), "-L", os.path.dirname(_lib.SO_PATH), "-lbgym",
                           "-Wl,-rpath," + os.path.dirname(_lib.SO_PATH)])
    out = subprocess.check_output([exe, "4096", "200"], text=True)
    assert "env-steps/s" in out and "episodes finished" in out
    episodes = int(out.split("env-steps/s,")[1].split("episodes")[0])
    assert episodes > 1000        # random play ends episodes all the time: the autoreset path ran
