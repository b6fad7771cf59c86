"""ctypes binding of the C oracle (oracle/bgym_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
Arrays are numpy (host) arrays of the C-ABI record dtypes (balatro_gym_b200/layout.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)
from balatro_gym_b200 import layout as L  # noqa: E402

_SO = os.path.join(_HERE, "libbgym_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bgym_oracle.c")
    deps = [src, os.path.join(_REPO, "include", "bgym.h"), os.path.join(_REPO, "include", "bgym_tables.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbgym_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        sizes = (C.c_int * 5)()
        _lib.oracle_sizes(sizes)
        assert list(sizes) == [L.STATE_BYTES, L.OBS_BYTES, L.INFO_BYTES, L.DRAWS_BYTES, 16], list(sizes)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def reset(state, obs, seeds, decks52=None, reset_mask=None, flags=0):
    n = state.shape[0]
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    if decks52 is not None:
        decks52 = np.ascontiguousarray(decks52, dtype=np.uint8)
        assert decks52.shape == (n, 52)
    if reset_mask is not None:
        reset_mask = np.ascontiguousarray(reset_mask, dtype=np.uint8)
    rc = lib().oracle_reset(_p(state), _p(obs), _p(reset_mask), _p(seeds), _p(decks52), C.c_int64(n), C.c_int(flags))
    assert rc == 0


def step(state, actions, obs, reward, terminated, truncated=None, info=None, draws=None, flags=0):
    n = state.shape[0]
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    rc = lib().oracle_step(_p(state), _p(actions), _p(draws), _p(obs), _p(reward), _p(terminated),
                           _p(truncated), _p(info), C.c_int64(n), C.c_int(flags))
    assert rc == 0


def action_mask(state):
    n = state.shape[0]
    out = np.zeros(n, dtype=np.uint64)
    lib().oracle_action_mask(_p(state), _p(out), C.c_int64(n))
    return out


def sample_actions(obs, seed, step_idx):
    n = obs.shape[0]
    out = np.zeros(n, dtype=np.int32)
    lib().oracle_sample_actions(_p(obs), _p(out), C.c_uint32(seed), C.c_uint64(step_idx), C.c_int64(n))
    return out


def score_hands(cards8, mods8=None, n_cards=None, jokers8=None, levels12=None, ctx=None, seed=0, flags=0):
    cards8 = np.ascontiguousarray(cards8, dtype=np.uint8)
    n = cards8.shape[0]
    assert cards8.shape == (n, 8)
    if mods8 is not None:
        mods8 = np.ascontiguousarray(mods8, dtype=np.uint16)
    if n_cards is not None:
        n_cards = np.ascontiguousarray(n_cards, dtype=np.uint8)
    if jokers8 is not None:
        jokers8 = np.ascontiguousarray(jokers8, dtype=np.uint8)
    if levels12 is not None:
        levels12 = np.ascontiguousarray(levels12, dtype=np.uint8)
    out = dict(hand_type=np.zeros(n, np.uint8), chips=np.zeros(n, np.int32), mult=np.zeros(n, np.int32),
               x_mult=np.zeros(n, np.float64), score=np.zeros(n, np.int64), money=np.zeros(n, np.int32))
    rc = lib().oracle_score_hands(_p(cards8), _p(mods8), _p(n_cards), _p(jokers8), _p(levels12), _p(ctx),
                                  _p(out["hand_type"]), _p(out["chips"]), _p(out["mult"]), _p(out["x_mult"]),
                                  _p(out["score"]), _p(out["money"]), C.c_uint32(seed), C.c_int64(n), C.c_int(flags))
    assert rc == 0
    return out


class OracleVec:
    """n envs stepped by the C oracle; same record arrays as the CUDA path."""

    def __init__(self, n):
        self.n = n
        self.state = np.zeros(n, dtype=L.STATE_DTYPE)
        self.obs = np.zeros(n, dtype=L.OBS_DTYPE)
        self.reward = np.zeros(n, dtype=np.float64)
        self.terminated = np.zeros(n, dtype=np.uint8)
        self.truncated = np.zeros(n, dtype=np.uint8)
        self.info = np.zeros(n, dtype=L.INFO_DTYPE)

    def reset(self, seeds, decks52=None, reset_mask=None):
        reset(self.state, self.obs, seeds, decks52, reset_mask)
        return self.obs

    def step(self, actions, draws=None, flags=0):
        step(self.state, actions, self.obs, self.reward, self.terminated, self.truncated, self.info, draws, flags)
        return self.obs, self.reward, self.terminated, self.truncated, self.info
