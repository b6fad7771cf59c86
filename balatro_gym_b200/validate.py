"""Validators — the reference's `BalatroEnvValidator` (balatro_env_2.py:1733-1796) for the Gymnasium
facade, plus batched versions that check the same two properties for every env of a device slab
at once, and a checkpoint round-trip check (save_state / load_state, :1575-1615).

The static methods keep the reference's names, arguments and failure behaviour (AssertionError with
the same messages), so they run unchanged against either implementation.
"""
from __future__ import annotations

import numpy as np

from . import layout as L


class BalatroEnvValidator:
    """Validate environment behaviour (duck-typed: works on the reference's BalatroEnv too)."""

    @staticmethod
    def validate_determinism(env_class, seed: int = 42, steps: int = 100):
        """Two envs built with the same seed, fed the same actions, must agree on everything (:1737)."""
        env1 = env_class(seed=seed)
        env2 = env_class(seed=seed)
        obs1, _ = env1.reset()
        obs2, _ = env2.reset()
        for key in obs1:
            if not np.array_equal(obs1[key], obs2[key]):
                raise AssertionError(f"Initial observations differ for key: {key}")
        for i in range(steps):
            valid_actions = np.where(obs1['action_mask'])[0]
            if len(valid_actions) == 0:
                break
            action = valid_actions[i % len(valid_actions)]
            obs1, r1, t1, tr1, _ = env1.step(action)
            obs2, r2, t2, tr2, _ = env2.step(action)
            if r1 != r2:
                raise AssertionError(f"Rewards differ at step {i}: {r1} vs {r2}")
            if t1 != t2 or tr1 != tr2:
                raise AssertionError(f"Termination differs at step {i}")
            for key in obs1:
                if not np.array_equal(obs1[key], obs2[key]):
                    raise AssertionError(f"Observations differ at step {i} for key: {key}")
        return True

    @staticmethod
    def validate_action_masking(env):
        """Masked actions are rejected with reward -1.0 and info['error']; legal ones are not (:1776).

        One deliberate difference: the reference's loop keeps testing against the mask of the RESET
        observation while the env moves on (action 45 leaves BLIND_SELECT, so "valid" action 46 is then
        rejected) and therefore fails on the reference itself; here the mask is re-read after every
        step, which is what the check means."""
        obs, _ = env.reset()
        for action in range(env.action_space.n):
            if obs['action_mask'][action]:
                obs, _, _, _, info = env.step(action)
                if 'error' in info and info['error'] == 'Invalid action':
                    raise AssertionError(f"Valid action {action} was rejected")
            else:
                obs, reward, _, _, info = env.step(action)
                if 'error' not in info:
                    raise AssertionError(f"Invalid action {action} was accepted")
                if reward != -1.0:
                    raise AssertionError(f"Invalid action {action} gave reward {reward}")
        return True


def _snapshot(vec):
    t = vec.torch
    return [x.clone() for x in (vec.hot, vec.cold, vec.obs_buf, vec.info_buf, vec.reward, vec.terminated)]


def validate_determinism_vec(num_envs: int = 4096, seed: int = 42, steps: int = 100, c3: bool = True, device="cuda"):
    """Batched determinism: two slabs with the same seeds stepped with the same (random legal) actions
    are BITWISE identical in state, observation, reward, termination and info after every step,
    across in-kernel autoresets."""
    from .vec_env import BalatroVecEnv
    a = BalatroVecEnv(num_envs, device=device, seed=seed)
    b = BalatroVecEnv(num_envs, device=device, seed=seed)
    for v in (a, b):
        v.reset()
        if c3:
            v.randomize_c3(seed)
    torch = a.torch
    for i in range(steps):
        acts = a.sample_actions(seed=seed)
        a.step(acts)
        b.step(acts.clone())
        for name, x, y in zip(("hot", "cold", "obs", "info", "reward", "terminated"), _snapshot(a), _snapshot(b)):
            if not torch.equal(x, y):
                raise AssertionError(f"{name} differs at step {i}")
    return True


def validate_action_masking_vec(num_envs: int = 4096, seed: int = 42, warm_steps: int = 40, device="cuda"):
    """Batched masking check from a mixed-phase state: for each of the 60 action ids, step a copy of
    the slab with that id everywhere; envs whose mask bit is clear must return reward -1.0,
    error_code INVALID_ACTION and an unchanged state record; envs whose bit is set must not report
    INVALID_ACTION.  The mask word from `bgym_action_mask` must equal the one inside the observation."""
    from .vec_env import BalatroVecEnv
    v = BalatroVecEnv(num_envs, device=device, seed=seed, autoreset=False)
    torch = v.torch
    v.reset()
    v.randomize_c3(seed)
    for _ in range(warm_steps):       # spread the slab over phases (terminated envs stay terminated: no autoreset)
        v.step(v.sample_actions(seed=seed))
    v.reset(reset_mask=v.terminated.clone())
    ck = v.save_state()
    bits = v.action_masks().clone()
    if not torch.equal(bits, v.obs["action_mask_bits"]):
        raise AssertionError("bgym_action_mask differs from the observation's mask word")
    err = v.info_field("error_code")
    for action in range(L.NUM_ACTIONS):
        v.load_state(ck)
        hot0, cold0 = v.hot.clone(), v.cold.clone()
        v.step(torch.full((num_envs,), action, dtype=torch.int32, device=v.device))
        legal = ((bits >> action) & 1).bool()
        bad = ~legal
        if bad.any():
            if not bool((v.reward[bad] == -1.0).all()):
                raise AssertionError(f"Invalid action {action} gave a reward other than -1.0")
            if not bool((err[bad] == L.ERR_INVALID_ACTION).all()):
                raise AssertionError(f"Invalid action {action} was accepted")
            # the record is unchanged apart from the step counter of the episode
            h1 = v.hot.clone()
            off = L.HOT_DTYPE.fields["ep_len"][1]
            h1[:, off:off + 4] = hot0[:, off:off + 4]
            if not (torch.equal(h1[bad], hot0[bad]) and torch.equal(v.cold[bad], cold0[bad])):
                raise AssertionError(f"Invalid action {action} changed the state")
        if legal.any() and bool((err[legal] == L.ERR_INVALID_ACTION).any()):
            raise AssertionError(f"Valid action {action} was rejected")
    return True


def validate_checkpoint_roundtrip(num_envs: int = 4096, seed: int = 3, steps: int = 50, device="cuda"):
    """save_state -> K steps -> load_state -> the same K steps reproduces every buffer bitwise
    (the reference's save_state/load_state contract, balatro_env_2.py:1575-1615, for a whole slab)."""
    from .vec_env import BalatroVecEnv
    v = BalatroVecEnv(num_envs, device=device, seed=seed)
    torch = v.torch
    v.reset()
    v.randomize_c3(seed)
    for _ in range(20):
        v.step(random_policy=True)
    ck = v.save_state()
    for _ in range(steps):
        v.step(random_policy=True)
    first = _snapshot(v)
    v.load_state(ck)
    for _ in range(steps):
        v.step(random_policy=True)
    for name, x, y in zip(("hot", "cold", "obs", "info", "reward", "terminated"), first, _snapshot(v)):
        if not torch.equal(x, y):
            raise AssertionError(f"{name} differs after the checkpoint round trip")
    return True
