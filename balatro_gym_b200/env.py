"""BalatroEnv — the Gymnasium facade (N = 1) over the device vector env.

Same public surface as the reference class `balatro_gym/balatro_env_2.py::BalatroEnv` (:354):
    BalatroEnv(*, render_mode=None, seed=None)          :359
    reset(*, seed=None, options=None) -> (obs, info)     :505
    step(action) -> (obs, reward, terminated, truncated, info)   :616
    action_space = Discrete(60)                          :368
    observation_space = Dict(...)                        :386-470 (the 31 keys actually emitted)
    save_state() / load_state()                          :1575 / :1595
    make_balatro_env(**kw) thunk                         :1803
Error convention as in the reference: a masked action returns reward -1.0 and
info['error'] without raising (:626-627).

Replaying the reference's shuffle: the reference shuffles with CPython's
`random.Random(seed % 2**32).shuffle` (stream 0 of DeterministicRNG, :84-106, :525).  With
`options={'shuffle': 'reference'}` (the default) the facade keeps that MT19937 stream: `reset(seed=s)`
restarts it, a plain `reset()` takes its NEXT shuffle — the constructor's own reset takes the first —
exactly as the reference does (:507-509), so `BalatroEnv(seed=s)` followed by any sequence of resets
sees the reference's decks.  `options={'shuffle': 'philox'}` uses the native counter-based stream.
"""
from __future__ import annotations

import random as _pyrandom
from typing import Optional

import numpy as np

from . import layout as L

try:  # real gymnasium if present, else the constructor-only stand-ins
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _EnvBase = _gym.Env
except Exception:  # pragma: no cover - this image has no gymnasium
    _gym = None
    from . import _spaces
    _EnvBase = object

_HAND_TYPE_NAMES = ["HIGH_CARD", "ONE_PAIR", "TWO_PAIR", "THREE_KIND", "STRAIGHT", "FLUSH", "FULL_HOUSE",
                    "FOUR_KIND", "STRAIGHT_FLUSH", "FIVE_KIND", "FLUSH_HOUSE", "FLUSH_FIVE"]
_ERR_TEXT = {L.ERR_INVALID_ACTION: "Invalid action", L.ERR_BOSS_RESTRICTION: "Boss blind restriction",
             L.ERR_CONSUMABLE_FAILED: "Failed to use consumable", L.ERR_SHOP: "Shop error",
             L.ERR_REF_EXCEPTION: "reference raises here (SafeBalatroEnv convention applied)",
             L.ERR_UNSUPPORTED: "unsupported consumable"}


def reference_deck(seed: int) -> np.ndarray:
    """The deck the reference builds for `seed`: suit-major/rank-minor order (:519-522) shuffled by
    `random.Random((seed + 0*1000) % 2**32).shuffle` (:105, :525).  Returns 52 card codes."""
    deck = [(rank - 2) * 4 + suit for suit in range(4) for rank in range(2, 15)]
    _pyrandom.Random(seed % (2 ** 32)).shuffle(deck)
    return np.asarray(deck, dtype=np.uint8)


def observation_space():
    S = _spaces
    return S.Dict({
        'hand': S.Box(-1, 51, (8,), dtype=np.int8), 'hand_size': S.Box(0, 12, (), dtype=np.int8),
        'deck_size': S.Box(0, 52, (), dtype=np.int8), 'selected_cards': S.MultiBinary(8),
        'chips_scored': S.Box(0, 10_000_000_000, (), dtype=np.int64),
        'round_chips_scored': S.Box(0, 10_000_000, (), dtype=np.int32),
        'progress_ratio': S.Box(0.0, 2.0, (), dtype=np.float32), 'mult': S.Box(0, 10_000, (), dtype=np.int32),
        'chips_needed': S.Box(0, 10_000_000, (), dtype=np.int32), 'money': S.Box(-20, 999, (), dtype=np.int32),
        'ante': S.Box(1, 1000, (), dtype=np.int16), 'round': S.Box(1, 3, (), dtype=np.int8),
        'hands_left': S.Box(0, 12, (), dtype=np.int8), 'discards_left': S.Box(0, 10, (), dtype=np.int8),
        'joker_count': S.Box(0, 10, (), dtype=np.int8), 'joker_ids': S.Box(0, 200, (10,), dtype=np.int16),
        'joker_slots': S.Box(0, 10, (), dtype=np.int8), 'consumable_count': S.Box(0, 5, (), dtype=np.int8),
        'consumables': S.Box(0, 100, (5,), dtype=np.int16), 'consumable_slots': S.Box(0, 5, (), dtype=np.int8),
        'shop_items': S.Box(0, 300, (10,), dtype=np.int16), 'shop_costs': S.Box(0, 5000, (10,), dtype=np.int16),
        'shop_rerolls': S.Box(0, 999, (), dtype=np.int16), 'hand_levels': S.Box(0, 15, (12,), dtype=np.int8),
        'phase': S.Box(0, 3, (), dtype=np.int8), 'action_mask': S.MultiBinary(L.NUM_ACTIONS),
        'hands_played': S.Box(0, 10000, (), dtype=np.int32),
        'best_hand_this_ante': S.Box(0, 10_000_000, (), dtype=np.int32),
        'boss_blind_active': S.Box(0, 1, (), dtype=np.int8), 'boss_blind_type': S.Box(0, 30, (), dtype=np.int8),
        'face_down_cards': S.MultiBinary(8),
    })


def info_dict(inf) -> dict:
    """Step-info dict from one BgymInfo record, with the reference's keys (balatro_env_2.py:626-627, 925-960)."""
    info = {}
    if inf["error_code"]:
        info["error"] = _ERR_TEXT.get(int(inf["error_code"]), "error")
    if inf["flags"] & L.F_PLAYED:
        info["final_score"] = int(inf["final_score"])
        info["hand_type"] = int(inf["hand_type"])
        info["hand_type_name"] = _HAND_TYPE_NAMES[int(inf["hand_type"])]
        info["cards_played"] = int(inf["cards_played"])
        info["score_breakdown"] = {"final_chips": int(inf["chips"]), "final_mult": int(inf["mult"]),
                                   "final_x_mult": float(inf["x_mult"]), "final_score": int(inf["base_score"])}
    if inf["flags"] & L.F_BEAT_BLIND:
        info["beat_blind"] = True
    if inf["flags"] & L.F_FAILED:
        info["failed"] = True
    if inf["flags"] & L.F_GUARD_TERMINATED:
        info["terminated"] = "guard"
    return info


class BalatroEnv(_EnvBase):
    """One env behind the Gymnasium protocol, driven through the C-ABI's host-buffer handle
    (`bgym_vec_create / reset_host / step_host`, include/bgym.h): per step ONE C call = action H2D, one kernel
    launch (small-slab step), one D2H of {obs, reward, info, flags} into numpy buffers.  No torch involved."""
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 4}

    def __init__(self, *, render_mode: Optional[str] = None, seed: Optional[int] = None, device=0):
        import ctypes as C
        from . import _lib
        self.render_mode = render_mode
        self._seed = seed if seed else int(np.random.randint(1, 2 ** 31 - 1))
        self.lib = _lib.load()
        if self.lib.bgym_device_count() < 1:
            raise _lib.BgymError("balatro_gym_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if isinstance(device, str):
            device = int(device.split(":")[1]) if ":" in device else 0
        elif not isinstance(device, int):        # torch.device
            device = device.index or 0
        self._h = C.c_void_p()
        _lib.check(self.lib.bgym_vec_create(C.byref(self._h), 1, int(device)), "bgym_vec_create")
        self._act = np.zeros(1, dtype=np.int32)
        self._seeds = np.zeros(1, dtype=np.uint32)
        self._obs_rec = np.zeros(1, dtype=L.OBS_DTYPE)
        self._rew = np.zeros(1, dtype=np.float64)
        self._term = np.zeros(1, dtype=np.uint8)
        self._trunc = np.zeros(1, dtype=np.uint8)
        self._info = np.zeros(1, dtype=L.INFO_DTYPE)
        self._state = np.zeros(1, dtype=L.STATE_DTYPE)
        self._p = {k: getattr(self, "_" + k).ctypes.data for k in ("act", "seeds", "obs_rec", "rew", "term", "trunc", "info", "state")}
        self.action_space = _spaces.Discrete(L.NUM_ACTIONS)
        self.observation_space = observation_space()
        self._new_streams(self._seed)
        self.reset()          # like the reference's constructor: consumes the first shuffle of the seed's stream

    # -- conversions ---------------------------------------------------------------------------------
    def _obs(self):
        rec = self._obs_rec[0]
        out = {}
        for k in L.OBS_KEYS:
            v = L.obs_value(rec, k)
            out[k] = v.copy() if isinstance(v, np.ndarray) else v
        return out

    def _new_streams(self, seed: int):
        """A seed (re)starts the env's random streams, as `DeterministicRNG(seed)` does in the reference
        (balatro_env_2.py:507-509): the MT19937 deck-shuffle stream (stream 0: `random.Random(seed % 2**32)`) and the
        Philox key of the in-game draws."""
        self._seed = int(seed)
        self._mt = _pyrandom.Random(self._seed % (2 ** 32))
        self._philox_seed = np.uint32(self._seed % (2 ** 32) or 1)

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """`reset(seed=s)` restarts the streams; `reset()` CONTINUES them, so every episode gets a new deck, boss and
        shops — the reference only rebuilds its RNG when a seed is passed (:507-509)."""
        from . import _lib
        from .sb3_vec_env import next_episode_seed
        if seed is not None and seed != 0:
            self._new_streams(seed)
        else:
            self._philox_seed = np.uint32(next_episode_seed(np.asarray([self._philox_seed]))[0])
        mode = (options or {}).get("shuffle", "reference")
        self._seeds[0] = self._philox_seed
        deck = None
        if mode == "reference":      # the next shuffle of the seed's MT19937 stream (:519-525)
            cards = [(rank - 2) * 4 + suit for suit in range(4) for rank in range(2, 15)]
            self._mt.shuffle(cards)
            deck = np.asarray(cards, dtype=np.uint8)
        rc = self.lib.bgym_vec_reset_host(self._h, self._p["seeds"], None if deck is None else deck.ctypes.data, self._p["obs_rec"])
        _lib.check(rc, "bgym_vec_reset_host")
        return self._obs(), {}

    def step(self, action: int):
        self._act[0] = int(action)
        p = self._p
        rc = self.lib.bgym_vec_step_host(self._h, p["act"], None, p["obs_rec"], p["rew"], p["term"], p["trunc"], p["info"], 0)
        if rc:
            from . import _lib
            _lib.check(rc, "bgym_vec_step_host")
        return self._obs(), float(self._rew[0]), bool(self._term[0]), False, info_dict(self._info[0])

    def action_masks(self):
        return L.mask_from_bits(self._obs_rec[0]["action_mask_bits"]).astype(bool)

    @property
    def state(self):
        """Host copy of the env's state record (numpy record of layout.STATE_DTYPE)."""
        from . import _lib
        _lib.check(self.lib.bgym_vec_get_state(self._h, self._p["state"]), "bgym_vec_get_state")
        return self._state[0].copy()

    def save_state(self):
        return {"state": self.state, "obs": self._obs_rec.copy()}

    def load_state(self, saved):
        from . import _lib
        self._state[0] = saved["state"]
        _lib.check(self.lib.bgym_vec_set_state(self._h, self._p["state"]), "bgym_vec_set_state")
        self._obs_rec[:] = saved["obs"]

    def render(self):
        if self.render_mode != "human":
            return
        s = self.state
        print(f"Ante {s['ante']} Round {s['round']} Phase {s['phase']} | score {s['round_chips']}/{s['chips_needed']} "
              f"| total {s['chips_scored']} | ${s['money']} | hands {s['hands_left']} discards {s['discards_left']}")

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.bgym_vec_destroy(h)
        return None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_balatro_env(**kwargs):
    """Factory thunk, as balatro_env_2.py:1803-1807."""
    def _init():
        return BalatroEnv(**kwargs)
    return _init


_REGISTRY = {"BalatroGym-v0": BalatroEnv}


def make(id: str, **kwargs):
    """`gym.make("BalatroGym-v0")` (README.md:37 of the reference, which never registers it)."""
    if _gym is not None:
        return _gym.make(id, **kwargs)
    return _REGISTRY[id](**kwargs)


def register_envs():
    if _gym is not None:
        try:
            _gym.register(id="BalatroGym-v0", entry_point="balatro_gym_b200.env:BalatroEnv")
        except Exception:
            pass
