"""Harness around the UNMODIFIED reference env (cassiusfive/balatro-gym) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (balatro_gym_b200/) never does.

What it provides
  load_reference()      import balatro_gym.balatro_env_2 from /root/reference (this container) or
                        from oracle/_ref (byte-compiled copy built by oracle/build_ref.py, which is
                        what travels to the GPU box), behind the gymnasium shim if needed.
  TapRandom             a random.Random that logs every draw semantically (SURVEY.md §7 "RNG tap").
  RefEnv                BalatroEnv + taps on all 16 DeterministicRNG streams, the Shop rng and the
                        module-global `random` used by boss_blinds / complete_joker_effects /
                        consumables; state injection; extraction of the env into the C-ABI records
                        (include/bgym.h) so the reference can be compared byte-for-byte with the
                        C oracle and the CUDA kernels.
"""
from __future__ import annotations

import os
import random
import sys
from typing import Any, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from balatro_gym_b200 import layout as L  # noqa: E402  (record dtypes = the ABI, not the product path)

REFERENCE_ROOT = "/root/reference"
REF_COMPILED = os.path.join(_HERE, "_ref")

_ref = None


def reference_available() -> Optional[str]:
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "balatro_gym")):
        return REFERENCE_ROOT
    if os.path.isfile(os.path.join(REF_COMPILED, "balatro_gym", "balatro_env_2.pyc")):
        return REF_COMPILED
    if os.path.isfile(os.path.join(REF_COMPILED, "balatro_gym_ref.zip")):
        return os.path.join(REF_COMPILED, "balatro_gym_ref.zip")
    return None


class _Ref:
    pass


class DeckCapacity(Exception):
    """The reference's deck list holds more Cryptid copies than BgymHot.deck_extra can (include/bgym.h)."""


def load_reference():
    """Import the reference package; returns a namespace with the modules we use."""
    global _ref
    if _ref is not None:
        return _ref
    root = reference_available()
    if root is None:
        raise RuntimeError("reference not available: neither /root/reference nor oracle/_ref "
                           "(run `python oracle/build_ref.py` where /root/reference is mounted)")
    try:
        import gymnasium  # noqa: F401
    except Exception:
        shim = os.path.join(_HERE, "refshim")
        if shim not in sys.path:
            sys.path.insert(0, shim)
    if root not in sys.path:
        sys.path.insert(0, root)
    sys.dont_write_bytecode = True  # never write into /root/reference
    import importlib
    ns = _Ref()
    ns.root = root
    ns.env_mod = importlib.import_module("balatro_gym.balatro_env_2")
    ns.boss = importlib.import_module("balatro_gym.boss_blinds")
    ns.jeff = importlib.import_module("balatro_gym.complete_joker_effects")
    ns.cons = importlib.import_module("balatro_gym.consumables")
    ns.shop = importlib.import_module("balatro_gym.shop")
    ns.cards = importlib.import_module("balatro_gym.cards")
    ns.jokers = importlib.import_module("balatro_gym.jokers")
    ns.scoring = importlib.import_module("balatro_gym.scoring_engine")
    ns.unified = importlib.import_module("balatro_gym.unified_scoring")
    ns.game = importlib.import_module("balatro_gym.balatro_game")
    ns.constants = importlib.import_module("balatro_gym.constants")
    ns.BalatroEnv = ns.env_mod.BalatroEnv
    ns.Action = ns.constants.Action
    ns.Phase = ns.constants.Phase
    ns.HandType = ns.scoring.HandType
    ns.JOKER_BY_ID = {j.id: j for j in ns.jokers.JOKER_LIBRARY}
    ns.JOKER_BY_NAME = {j.name: j for j in ns.jokers.JOKER_LIBRARY}
    _ref = ns
    return ns


_rules = None


def load_rules_evaluator():
    """The UNMODIFIED rules evaluator, `balatro_gym/balatro_sim.py::BalatroSimulator` (its `evaluate_hand`,
    :220-400, pins BGYM_SCORE_RULES).  The module imports `scoring_engine` as a top-level name (:6), so the
    package directory itself goes on sys.path as well."""
    global _rules
    if _rules is None:
        R = load_reference()
        pkg = os.path.join(R.root, "balatro_gym")      # a directory, or a path inside the zip archive (zipimport takes both)
        if pkg not in sys.path:
            sys.path.append(pkg)
        import importlib
        _rules = importlib.import_module("balatro_gym.balatro_sim")
    return _rules


# ---------------------------------------------------------------------------------------------
# RNG tap
# ---------------------------------------------------------------------------------------------
class TapRandom(random.Random):
    """random.Random that produces the identical stream and logs each draw semantically.

    log entries: ('u', float)  random()/uniform
                 ('k', int)    0-based index of a randint/choice/sample element
                 ('perm', list) shuffle result as the applied permutation
    CPython 3.12 internals relied on: choice/randint/sample/shuffle go through _randbelow
    (getrandbits), uniform goes through random().
    """

    def __new__(cls, seed=None, log=None):
        return super().__new__(cls, seed)

    def __init__(self, seed=None, log=None):
        super().__init__(seed)
        self.log = log if log is not None else []

    def Random(self, seed=None):  # lets an instance stand in for the `random` module (shop.py:99)
        return TapRandom(seed, self.log)

    def getrandbits(self, k):
        # Defining getrandbits keeps Random.__init_subclass__ from switching _randbelow to the
        # random()-based variant (it does that for subclasses that override random() only), so the
        # integer draws consume the Mersenne Twister exactly as a plain random.Random does.
        return super().getrandbits(k)

    def random(self):
        v = super().random()
        self.log.append(("u", v))
        return v

    def choice(self, seq):
        if not len(seq):
            raise IndexError("Cannot choose from an empty sequence")
        i = self._randbelow(len(seq))
        self.log.append(("k", i))
        return seq[i]

    def randint(self, a, b):
        v = a + self._randbelow(b - a + 1)
        # the one wide randint on the path is the Shop seed (balatro_env_2.py:1389); it only seeds
        # another tapped generator, so it is logged but never consumed by the kernels
        self.log.append(("k", v - a) if b - a < 256 else ("seed", v))
        return v

    def sample(self, population, k, *, counts=None):
        assert counts is None
        idx = super().sample(range(len(population)), k)
        for i in idx:
            self.log.append(("k", i))
        return [population[i] for i in idx]

    def shuffle(self, x):
        tmp = list(range(len(x)))
        super().shuffle(tmp)
        x[:] = [x[i] for i in tmp]
        self.log.append(("perm", tmp))


def draws_record(log) -> np.ndarray:
    """Pack a step's tap log into one BgymDraws record."""
    rec = np.zeros((), dtype=L.DRAWS_DTYPE)
    us = [v for t, v in log if t == "u"]
    ks = [v for t, v in log if t == "k"]
    if len(us) > 24 or len(ks) > 32:
        raise OverflowError(f"tap log too long for BgymDraws: {len(us)} uniforms, {len(ks)} ints")
    rec["u"][:len(us)] = us
    rec["k"][:len(ks)] = ks
    rec["n_u"] = len(us)
    rec["n_k"] = len(ks)
    return rec


# ---------------------------------------------------------------------------------------------
# RefEnv
# ---------------------------------------------------------------------------------------------
_TITLE_NAMES = ['High Card', 'One Pair', 'Two Pair', 'Three Kind', 'Straight', 'Flush', 'Full House',
                'Four Kind', 'Straight Flush', 'Five Kind', 'Flush House', 'Flush Five']


def card_code(card) -> int:
    return (int(card.rank) - 2) * 4 + int(card.suit)


class RefEnv:
    """The reference BalatroEnv with every RNG tapped, plus inject/extract helpers."""

    def __init__(self, seed: int, global_seed: int = 12345, tap: bool = True):
        assert seed >= 1, "seed 0/None is replaced by a random seed in the reference (SURVEY Q1)"
        self.R = load_reference()
        self.log: List[Any] = []
        self.tap = tap
        self.global_rng = TapRandom(global_seed, self.log)
        if tap:
            # module-global `random` of the three modules that use it, and shop.random.Random
            self.R.boss.random = self.global_rng
            self.R.jeff.random = self.global_rng
            self.R.cons.random = self.global_rng
            self.R.shop.random = self.global_rng
        self.env = self.R.BalatroEnv(seed=seed)
        self.seed = seed
        self.reset(seed)

    # -- lifecycle ---------------------------------------------------------------------------
    def reset(self, seed: int):
        self.seed = seed
        self.log.clear()
        if self.tap:
            self.R.boss.random = self.global_rng
            self.R.jeff.random = self.global_rng
            self.R.cons.random = self.global_rng
            self.R.shop.random = self.global_rng
        # reset(seed=) rebuilds DeterministicRNG (balatro_env_2.py:507-509); we need the taps in
        # place BEFORE the shuffle, so rebuild it ourselves exactly as :93-106 does.
        env = self.env
        env._seed = seed
        env.rng = self.R.env_mod.DeterministicRNG(seed)
        if self.tap:
            for i, name in enumerate(list(env.rng.streams.keys())):
                env.rng.streams[name] = TapRandom((seed + i * 1000) % (2 ** 32), self.log)
        # The reference keeps the previous Shop object across resets (reset() never touches
        # self.shop); it is unobservable (shop obs/mask are read only in SHOP phase, entered only
        # through _generate_shop which builds a new Shop), so the harness drops it to make the
        # extracted shop block well defined.
        env.shop = None
        obs, info = env.reset()
        self.deck_perm = None
        for t, v in self.log:
            if t == "perm":
                self.deck_perm = v
        self.log.clear()
        return obs, info

    def step(self, action: int):
        self.log.clear()
        out = self.env.step(int(action))
        return out

    def step_draws(self) -> np.ndarray:
        return draws_record(self.log)

    # -- views -------------------------------------------------------------------------------
    def deck_codes(self) -> np.ndarray:
        return np.array([card_code(c) for c in self.env.state.deck], dtype=np.uint8)

    def legal_actions(self) -> np.ndarray:
        return np.flatnonzero(self.env._get_action_mask())

    # -- injection -----------------------------------------------------------
    def inject_jokers(self, ids):
        self.env.state.jokers = [self.R.JOKER_BY_ID[i] for i in ids]

    def inject_card_mod(self, deck_idx, enhancement=0, edition=0, seal=0):
        C = self.R.cards
        self.env.state.card_states[deck_idx] = C.CardState(
            deck_idx, C.Enhancement(enhancement), C.Edition(edition), C.Seal(seal))

    def inject_hand_level(self, hand_type: int, level: int):
        ht = self.R.HandType(hand_type)
        self.env.engine.hand_levels[ht] = min(level, 15)
        self.env.state.hand_levels[ht] = level

    def inject_consumables(self, names):
        self.env.state.consumables = list(names)

    def force_boss(self, boss_type: Optional[int]):
        """Make select_boss_blind return a fixed type (None = restore tapped random.choice)."""
        if boss_type is None:
            self.R.env_mod.select_boss_blind = self.R.boss.select_boss_blind
        else:
            bt = self.R.boss.BossBlindType(boss_type)
            self.R.env_mod.select_boss_blind = lambda ante, exclude=None: bt

    # -- extraction into the C-ABI records ---------------------------------------------------------
    def extract_state(self) -> np.ndarray:
        env, R = self.env, self.R
        st = env.state
        s = np.zeros((), dtype=L.STATE_DTYPE)
        deck = st.deck
        hand = list(st.hand_indexes)
        assert len(hand) <= 8
        s["hand"][:] = 0xFF
        s["hand_code"][:] = 0xFF
        for i, idx in enumerate(hand):
            s["hand"][i] = idx
            if idx < len(deck):
                s["hand_code"][i] = card_code(deck[idx])
        s["hand_n"] = len(hand)
        s["hand_size"] = st.hand_size
        sel = list(st.selected_cards)
        s["sel_n"] = len(sel)
        so = 0
        for k, slot in enumerate(sel):
            so |= (slot & 15) << (4 * k)
        s["sel_order"] = so
        hm = 0
        for slot in env.game.highlighted_indexes:
            hm |= 1 << slot
        s["highlight_mask"] = hm
        fm = 0
        for slot in st.face_down_cards:
            if slot < 8:
                fm |= 1 << slot
        s["face_down_mask"] = fm
        s["phase"] = int(st.phase)
        s["round"] = st.round
        mgr = env.boss_blind_manager
        s["boss_type"] = int(st.active_boss_blind) if st.active_boss_blind else 0
        assert bool(st.boss_blind_active) == (mgr.active_blind is not None) == bool(st.active_boss_blind)
        s["hands_left"] = st.hands_left
        s["discards_left"] = st.discards_left
        s["joker_n"] = len(st.jokers)
        s["cons_n"] = len(st.consumables)
        s["joker_slots"] = st.joker_slots
        s["cons_slots"] = st.consumable_slots
        s["n_magic_trick"] = st.vouchers.count("Magic Trick")
        s["n_minimalist"] = st.vouchers.count("Minimalist")
        s["ante"] = st.ante
        s["jokers_sold"] = st.jokers_sold
        s["money"] = st.money
        s["chips_needed"] = st.chips_needed
        s["round_chips"] = st.round_chips_scored
        s["chips_scored"] = st.chips_scored
        s["best_hand"] = min(st.best_hand_this_ante, 2 ** 31 - 1)
        s["hands_played_total"] = st.hands_played_total
        s["hands_played_ante"] = st.hands_played_ante
        bs = mgr.blind_state if mgr.active_blind is not None else {}
        if bs:
            s["boss_flags"] = 1 if bs.get("first_hand") else 0
            s["boss_cards_required"] = bs.get("cards_required", 0)
            m = 0
            for name in bs.get("played_hand_types", ()):
                m |= 1 << _TITLE_NAMES.index(name)
            s["boss_played_types"] = m
            s["boss_hands_played"] = bs.get("hands_played", 0)
            ids = {id(c): i for i, c in enumerate(deck)}
            pm = 0
            for cid in bs.get("played_cards", ()):
                if cid in ids:             # a played card Immolate destroyed since is no longer in the deck
                    pm |= 1 << ids[cid]
            s["boss_played_cards"] = pm
        s["deck_n"] = len(deck)
        for i, j in enumerate(st.jokers):
            s["joker_id"][i] = j.id
        for i, name in enumerate(st.consumables):
            s["cons_id"][i] = L.consumable_id(name)
        for ht in R.HandType:
            s["hand_level"][int(ht)] = st.hand_levels.get(ht, 0)
            assert env.engine.hand_levels[ht] == min(15, max(1, st.hand_levels.get(ht, 1))), \
                "engine/state hand levels diverged beyond the min(level,15) relation"
            s["hand_play_count"][int(ht)] = min(255, env.engine.hand_play_counts[ht])
        s["shop_reroll_state"] = st.shop_reroll_cost
        # deck[i] = code of the card now at list index i (0 beyond the list) | card_states[i], which the reference keys
        # by INDEX; cards appended by Cryptid (consumables.Card objects, always a suffix of the list) are also listed
        # in deck_extra with the modifiers they were created with (their dataclass equality, include/bgym.h)
        for i in range(52):
            code = card_code(deck[i]) if i < len(deck) else 0
            cs = st.card_states.get(i)
            if cs is not None:
                s["deck"][i] = L.card16(code, int(cs.enhancement), int(cs.edition), int(cs.seal))
            else:
                s["deck"][i] = L.card16(code)
        extra = [c for c in deck if isinstance(c, R.cons.Card)]
        if len(extra) > 4:
            raise DeckCapacity(len(extra))
        assert all(isinstance(c, R.cons.Card) for c in deck[len(deck) - len(extra):])
        for j, c in enumerate(extra):
            s["deck_extra"][j] = L.card16(card_code(c), int(c.enhancement), int(c.edition), int(c.seal))
        s["deck_extra_n"] = len(extra)
        shop = env.shop
        if shop is not None:
            inv = shop.inventory
            assert len(inv) <= 9
            s["n_items"] = len(inv)
            for i, it in enumerate(inv):
                s["item_type"][i] = int(it.item_type)
                s["item_cost"][i] = it.cost
                p = it.payload
                if "joker_id" in p:
                    s["item_id"][i] = p["joker_id"]
                elif "pack_type" in p:
                    s["item_id"][i] = ["Standard Pack", "Joker Pack", "Tarot Pack", "Planet Pack",
                                       "Spectral Pack"].index(p["pack_type"])
                elif "voucher" in p:
                    s["item_id"][i] = ["Magic Trick", "Minimalist"].index(p["voucher"])
                else:
                    s["item_id"][i] = p["card"]
            s["reroll_cost"] = shop.reroll_cost
        return s

    @staticmethod
    def obs_record(obs: dict) -> np.ndarray:
        o = np.zeros((), dtype=L.OBS_DTYPE)
        for k in L.OBS_KEYS:
            if k != "action_mask":       # carried as action_mask_bits (include/bgym.h)
                o[k] = obs[k]
        bits = 0
        for a in np.flatnonzero(obs["action_mask"]):
            bits |= 1 << int(a)
        o["action_mask_bits"] = bits
        return o


# fields of BgymState that have no counterpart in the reference (native RNG bookkeeping)
STATE_NOCOMPARE = ("rng_seed", "rng_ctr", "ep_len", "episode")


def state_diff(a: np.ndarray, b: np.ndarray, skip=STATE_NOCOMPARE):
    """Names of fields that differ between two BgymState records."""
    out = []
    for name in L.STATE_DTYPE.names:
        if name in skip:
            continue
        if not np.array_equal(a[name], b[name]):
            out.append((name, a[name].tolist(), b[name].tolist()))
    return out


def obs_diff(a: np.ndarray, b: np.ndarray):
    out = []
    for name in L.OBS_DTYPE.names:
        if not np.array_equal(a[name], b[name]):
            out.append((name, a[name].tolist(), b[name].tolist()))
    return out
