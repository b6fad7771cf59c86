"""BalatroSB3VecEnv — the Stable-Baselines3 `VecEnv` protocol over the device vector env.

The reference trains through SB3 vector envs built from per-process `BalatroEnv` instances
(SURVEY §8(f)1):
    hpc_train.py:57-72            SubprocVecEnv([make_env(i, seed) ...]) / DummyVecEnv -> VecNormalize -> PPO
    train_balatro_fixed.py:285-330   BalatroEnvFixed -> SafeBalatroEnv(max_invalid_actions=50,
                                     max_episode_steps=1000) -> Monitor -> SubprocVecEnv
This class is the drop-in for that stack: one object, `num_envs` envs stepped by the sm_100a
kernels, with the SB3 conventions the trainers rely on

  * `reset() -> obs`, `step_async(actions)`, `step_wait() -> (obs, rewards, dones, infos)`;
    observations are a dict of numpy arrays stacked on axis 0 (the 31 keys of
    balatro_env_2.py:1488-1531, same dtypes);
  * same-step auto-reset as `DummyVecEnv.step_wait`: a finished env returns the FIRST observation
    of its next episode and `infos[i]['terminal_observation']` holds the last one of the old;
  * `Monitor`'s `infos[i]['episode'] = {'r', 'l', 't'}` on episode end;
  * `SafeBalatroEnv`'s guards (train_balatro_fixed.py:240-258): `max_invalid_actions` consecutive
    rejected actions (reward == -1.0) end the episode with reward -50 and
    `info['invalid_action_termination']`; `max_episode_steps` sets `TimeLimit.truncated` /
    `info['max_steps_reached']`.  Both are off (None) by default like the bare env;
  * `env_method('action_masks')` for sb3-contrib's MaskablePPO, `get_attr / set_attr / seed`.

It subclasses `stable_baselines3.common.vec_env.VecEnv` when SB3 is importable and is a duck-typed
stand-in otherwise (this image has no SB3).  It drives the C-ABI's host-buffer handle (`bgym_vec_*`,
numpy buffers, no torch): a step is ONE C call = actions H2D, the step kernels, ONE device->host copy
of {observations, rewards, infos, flags}; what the adapter adds is host-side bookkeeping on
[num_envs] numpy arrays.  For throughput at scale use `BalatroVecEnv` / `RolloutCollector`, which
never leave the device.
"""
from __future__ import annotations

import random as _pyrandom
import time
from typing import Any, List, Optional, Sequence

import numpy as np

from . import layout as L
from .env import observation_space as _observation_space, info_dict

try:  # pragma: no cover - not in this image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
    _HAVE_SB3 = True
except Exception:
    _VecEnvBase = object
    _HAVE_SB3 = False

try:
    from gymnasium import spaces as _spaces  # pragma: no cover
except Exception:
    from . import _spaces


def next_episode_seed(seed: np.ndarray) -> np.ndarray:
    """Host mirror of the kernels' episode-seed chain (csrc/bgym_env.cuh next_episode_seed)."""
    x = (np.asarray(seed, dtype=np.uint64) + np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13); x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return np.where(x == 0, np.uint64(1), x).astype(np.uint32)


class BalatroSB3VecEnv(_VecEnvBase):
    """SB3-protocol vector env.

    shuffle : 'philox' (native counter-based decks, seed chain as in the kernels' autoreset) or
              'reference' (replay of the reference's MT19937 deck stream: env i behaves like
              `BalatroEnv(seed=seed+i)` driven by `DummyVecEnv`, i.e. its k-th `reset()` uses the
              (k+1)-th shuffle of `random.Random(seed+i)` — the constructor's own reset consumed
              the first, balatro_env_2.py:384,505-525).
    """

    def __init__(self, num_envs: int, seed: int = 1, device=0, shuffle: str = "philox",
                 max_invalid_actions: Optional[int] = None, max_episode_steps: Optional[int] = None,
                 monitor: bool = True):
        import ctypes as C
        from . import _lib
        assert shuffle in ("philox", "reference")
        self._libmod = _lib
        self.lib = _lib.load()
        if self.lib.bgym_device_count() < 1:
            raise _lib.BgymError("balatro_gym_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if isinstance(device, str):
            device = int(device.split(":")[1]) if ":" in device else 0
        elif not isinstance(device, int):
            device = device.index or 0
        self._h = C.c_void_p()
        _lib.check(self.lib.bgym_vec_create(C.byref(self._h), int(num_envs), int(device)), "bgym_vec_create")
        self.shuffle = shuffle
        self.max_invalid_actions = max_invalid_actions
        self.max_episode_steps = max_episode_steps
        self.monitor = monitor
        obs_space = _observation_space()
        act_space = _spaces.Discrete(L.NUM_ACTIONS)
        if _HAVE_SB3:  # pragma: no cover
            super().__init__(num_envs, obs_space, act_space)
        else:
            self.num_envs = num_envs
            self.observation_space = obs_space
            self.action_space = act_space
            self.reset_infos: List[dict] = [{} for _ in range(num_envs)]
        self.render_mode = None
        n = num_envs
        # host buffers the C-ABI fills (one device->host copy per step, include/bgym.h bgym_vec_step_host)
        self._rec = np.zeros(n, dtype=L.OBS_DTYPE)
        self._reward = np.zeros(n, dtype=np.float64)
        self._term = np.zeros(n, dtype=np.uint8)
        self._trunc = np.zeros(n, dtype=np.uint8)
        self._info = np.zeros(n, dtype=L.INFO_DTYPE)
        self._actions = np.zeros(n, dtype=np.int32)
        self._state = np.zeros(n, dtype=L.STATE_DTYPE)
        self._seeds = (np.arange(n, dtype=np.int64) + seed) % (2 ** 32)
        self._seeds[self._seeds == 0] = 1
        self._seeds = self._seeds.astype(np.uint32)
        self._mt: Optional[List[_pyrandom.Random]] = None
        self._ep_ret = np.zeros(n, dtype=np.float64)
        self._ep_len = np.zeros(n, dtype=np.int64)
        self._invalid_run = np.zeros(n, dtype=np.int64)
        self._t0 = time.time()
        self._pending = None

    # -- observation plumbing ----------------------------------------------------------------------
    def _obs_dict(self, rec: np.ndarray) -> dict:
        """Fresh arrays, as `_get_observation` allocates them (balatro_env_2.py:1488)."""
        return {k: np.array(L.obs_value(rec, k)) for k in L.OBS_KEYS}

    def _host_records(self) -> np.ndarray:
        return self._rec

    def _reference_decks(self, idx: Sequence[int]) -> np.ndarray:
        decks = np.empty((len(idx), 52), dtype=np.uint8)
        for r, i in enumerate(idx):
            deck = [(rank - 2) * 4 + suit for suit in range(4) for rank in range(2, 15)]
            self._mt[i].shuffle(deck)
            decks[r] = deck
        return decks

    def _device_reset(self, mask: Optional[np.ndarray]):
        """Reset the envs selected by mask (all if None) with their current seeds / next decks; the
        observation records of ALL envs come back in self._rec."""
        n = self.num_envs
        decks = None
        if self.shuffle == "reference":
            idx = np.arange(n) if mask is None else np.flatnonzero(mask)
            decks = np.zeros((n, 52), dtype=np.uint8)
            decks[idx] = self._reference_decks(idx)
        m8 = None if mask is None else np.ascontiguousarray(mask.astype(np.uint8))
        seeds = np.ascontiguousarray(self._seeds)
        rc = self.lib.bgym_vec_reset_masked_host(self._h, None if m8 is None else m8.ctypes.data, seeds.ctypes.data,
                                                 None if decks is None else decks.ctypes.data, self._rec.ctypes.data)
        self._libmod.check(rc, "bgym_vec_reset_masked_host")

    # -- VecEnv protocol -----------------------------------------------------------------------------
    def seed(self, seed: Optional[int] = None):
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        s = (np.arange(self.num_envs, dtype=np.int64) + seed) % (2 ** 32)
        s[s == 0] = 1
        self._seeds = s.astype(np.uint32)
        self._mt = None
        return [int(x) for x in self._seeds]

    def reset(self):
        if self.shuffle == "reference" and self._mt is None:
            self._mt = [_pyrandom.Random(int(s)) for s in self._seeds]
            for i in range(self.num_envs):   # the reference constructor's own reset() (:384)
                self._mt[i].shuffle(list(range(52)))
        self._device_reset(None)
        self._ep_ret[:] = 0; self._ep_len[:] = 0; self._invalid_run[:] = 0
        self.reset_infos = [{} for _ in range(self.num_envs)]
        return self._obs_dict(self._rec)

    def step_async(self, actions):
        self._pending = np.asarray(actions).astype(np.int32).reshape(self.num_envs)

    def step_wait(self):
        self._actions[:] = self._pending
        rc = self.lib.bgym_vec_step_host(self._h, self._actions.ctypes.data, None, self._rec.ctypes.data, self._reward.ctypes.data,
                                         self._term.ctypes.data, self._trunc.ctypes.data, self._info.ctypes.data, 0)
        if rc:
            self._libmod.check(rc, "bgym_vec_step_host")
        rewards = self._reward.copy()
        terminated = self._term.astype(bool)
        info_rec = self._info
        truncated = np.zeros(self.num_envs, dtype=bool)
        # most steps carry nothing in info (a card toggle): only records with an error or flags become dicts
        infos: List[dict] = [{} for _ in range(self.num_envs)]
        for i in np.flatnonzero((info_rec["error_code"] != 0) | (info_rec["flags"] != 0)):
            infos[i] = info_dict(info_rec[i])

        # SafeBalatroEnv guards (train_balatro_fixed.py:240-258)
        self._ep_len += 1
        if self.max_invalid_actions is not None:
            rejected = (rewards == -1.0) & ~terminated
            self._invalid_run = np.where(rejected, self._invalid_run + 1, 0)
            hit = self._invalid_run >= self.max_invalid_actions
            for i in np.flatnonzero(hit):
                infos[i]['invalid_action_termination'] = True
            rewards = np.where(hit, -50.0, rewards)
            terminated = terminated | hit
        if self.max_episode_steps is not None:
            over = self._ep_len >= self.max_episode_steps
            for i in np.flatnonzero(over):
                infos[i]['max_steps_reached'] = True
            truncated = over
        self._ep_ret += rewards
        dones = terminated | truncated
        obs = self._obs_dict(self._rec)

        if dones.any():
            idx = np.flatnonzero(dones)
            for i in idx:
                infos[i]['terminal_observation'] = {k: obs[k][i].copy() for k in L.OBS_KEYS}
                infos[i]['TimeLimit.truncated'] = bool(truncated[i] and not terminated[i])
                if self.monitor:
                    infos[i]['episode'] = {'r': float(round(self._ep_ret[i], 6)), 'l': int(self._ep_len[i]),
                                           't': round(time.time() - self._t0, 6)}
            # every new episode gets a new key for its in-game draws (boss pick, shop inventories, glass / lucky
            # rolls), whichever way its deck is shuffled — with 'reference' decks the reference's own streams go on
            # from where the last episode left them, they do not restart either (balatro_env_2.py:507-509)
            self._seeds[idx] = next_episode_seed(self._seeds[idx])
            self._device_reset(dones)
            for k in L.OBS_KEYS:
                obs[k][idx] = L.obs_value(self._rec[idx], k)
            self._ep_ret[idx] = 0; self._ep_len[idx] = 0; self._invalid_run[idx] = 0
            for i in idx:
                self.reset_infos[i] = {}
        return obs, rewards.astype(np.float32), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.bgym_vec_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def action_masks(self) -> np.ndarray:
        """[num_envs, 60] bool, from the observation records already on the host."""
        return L.mask_from_bits(self._rec['action_mask_bits']).astype(bool)

    def env_method(self, method_name: str, *args, indices=None, **kwargs):
        idx = self._indices(indices)
        if method_name == "action_masks":
            m = self.action_masks()
            return [m[i] for i in idx]
        raise AttributeError(f"env_method({method_name!r}) is not available on the device vector env")

    def _pull_state(self) -> np.ndarray:
        self._libmod.check(self.lib.bgym_vec_get_state(self._h, self._state.ctypes.data), "bgym_vec_get_state")
        return self._state

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        idx = self._indices(indices)
        if attr_name == "render_mode":
            return [None for _ in idx]
        if attr_name in L.HOT_FIELD_NAMES or attr_name in L.COLD_FIELD_NAMES:
            col = self._pull_state()[attr_name]
            return [col[i].copy() if isinstance(col[i], np.ndarray) else col[i] for i in idx]
        if hasattr(self, attr_name):
            return [getattr(self, attr_name) for _ in idx]
        raise AttributeError(attr_name)

    def set_attr(self, attr_name: str, value, indices=None):
        idx = self._indices(indices)
        if attr_name in L.HOT_FIELD_NAMES or attr_name in L.COLD_FIELD_NAMES:
            st = self._pull_state()
            st[attr_name][idx] = value
            self._libmod.check(self.lib.bgym_vec_set_state(self._h, st.ctypes.data), "bgym_vec_set_state")
            return
        setattr(self, attr_name, value)

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode: Optional[str] = None):
        return None

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, int):
            return [indices]
        return list(indices)


def make_sb3_vec_env(n_envs: int, seed: int = 0, **kwargs) -> BalatroSB3VecEnv:
    """What `SubprocVecEnv([make_env(i, seed) for i in range(n_envs)])` builds in hpc_train.py:60."""
    return BalatroSB3VecEnv(n_envs, seed=seed, **kwargs)
