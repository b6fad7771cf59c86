#!/usr/bin/env python
"""Record golden traces from the UNMODIFIED reference (needs /root/reference or oracle/_ref).

    python tests/golden/make_golden.py

For each config a batch of episodes is played on the reference with random legal actions (plus a few
masked ones); every step stores the action, the reference's own random draws (BgymDraws record),
and the reference's results restated as C-ABI records: state after the step, observation, reward,
terminated, info numerics.  The traces are laid out [T, E] (step-major) so E episodes replay in
parallel on the GPU; after an episode ends its column is padded with action -1 (masked no-op).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
from oracle.refenv import RefEnv, DeckCapacity  # noqa: E402
from balatro_gym_b200 import layout as L  # noqa: E402
from lockstep import inject_c3  # noqa: E402


def c4x_plans():
    """Scripted episode openings of the c4x trace: every consumable name the reference knows (consumables.py:87-110,
    344-362 + the planets, balatro_env_2.py:1103-1116), in its stored form AND — for tarots — in the enum-style form
    The Emperor creates (consumables.py:172), each once without and once with selected target cards; then the
    deck-rebuilding pairs (Cryptid / Immolate / The Fool) incl. Immolate under The Pillar (played_cards follow the cards).
    A plan = (consumable names, number of target cards to select first, first action, boss to wait for or None,
    hands to play before the consumable is used, jokers to keep of the five injected ones)."""
    names = L.TAROT_NAMES + L.PLANET_NAMES + L.SPECTRAL_NAMES + [t.upper().replace(' ', '_') for t in L.TAROT_NAMES]
    plans = []
    for i, nm in enumerate(names):
        plans.append(([nm], 0, 45, None, 0, 5))
        plans.append(([nm], 1 + i % 3, 45 + i % 3, None, 0, 5))
    for nm in ('Wraith', 'The Soul', 'Ankh', 'Hex', 'Ectoplasm', 'Temperance'):   # joker-creating / joker-counting ones
        for keep in (0, 3, 4):                                                    # with free joker slots / no jokers
            plans.append(([nm, nm], 1, 46, None, 0, keep))
    combos = [['Cryptid', 'Immolate'], ['Immolate', 'Immolate'], ['Cryptid', 'Cryptid', 'Immolate', 'Immolate'],
              ['The Fool', 'Immolate'], ['Cryptid', 'The Fool', 'Immolate'], ['Immolate', 'Cryptid', 'Cryptid'],
              ['Cryptid', 'Immolate', 'Cryptid', 'Immolate', 'The Fool'], ['Immolate', 'Immolate', 'Immolate', 'Immolate'],
              ['The Emperor', 'The Fool'], ['The High Priestess', 'Judgement'], ['Wraith', 'Ectoplasm', 'Ouija']]
    for j, c in enumerate(combos):
        for rep in range(3):
            plans.append((c, 1 + (j + rep) % 2, 45 + rep, None, 0, 5))
    for rep in range(6):     # Immolate after hands were played under The Pillar (boss 16)
        plans.append((['Immolate', 'Cryptid', 'Immolate'][:1 + rep % 3], 1, 47, 16, 1 + rep % 2, 5))
    return plans


class _FakeOV:  # inject_c3 writes the same mods into an oracle-side record; here we only need the ref side
    def __init__(self):
        self.state = np.zeros(1, dtype=L.STATE_DTYPE)


def record(config, episodes, seed0, max_steps, invalid_rate=0.02):
    rng = np.random.default_rng(seed0)
    ref = RefEnv(seed=1)
    eps = []
    plans = c4x_plans() if config == "c4x" else None
    if plans is not None:
        episodes = max(episodes, len(plans))
    ep = 0
    attempt = 0
    while ep < episodes:
        seed = seed0 + ep + 7919 * attempt
        obs, _ = ref.reset(seed)
        deck = ref.deck_codes()
        if config in ("c3", "c4", "c4x"):
            fake = _FakeOV()
            fake.state[0]["deck"][:] = deck
            inject_c3(ref, fake, rng)
        script, want_boss = [], None
        if plans is not None:
            names, n_tgt, first_action, want_boss, n_plays, keep_jokers = plans[ep % len(plans)]
            ref.inject_consumables(names)
            ref.env.state.jokers = ref.env.state.jokers[:keep_jokers]
            script = [first_action]
            for _ in range(n_plays):               # play a few hands first (under a boss: fills played_cards)
                script += [2 + int(x) for x in rng.choice(8, size=2, replace=False)] + [0]
            script += [2 + int(x) for x in rng.choice(8, size=n_tgt, replace=False)] + [10]
        init_state = ref.extract_state()
        steps = []
        first = True
        retry = False
        for t in range(max_steps):
            legal = ref.legal_actions()
            if t < len(script):
                a = script[t]
            elif config == "c1" and first:
                a = 45
            elif config == "c3" and first:
                a = 47
            else:
                a = int(rng.choice(legal))
                if rng.random() < invalid_rate:
                    a = int(rng.integers(0, 60))
            first = False
            try:
                obs, r, term, trunc, info = ref.step(a)
                state = ref.extract_state()
            except OverflowError:
                break  # numpy-2-only int16 overflow of the reference's obs (pinned numpy 1.26 wraps)
            except DeckCapacity:
                break  # more than 4 appended cards: outside what BgymHot.deck_extra holds (include/bgym.h)
            except Exception:
                # reference raises: expected -100 / terminated / state unchanged
                steps.append(dict(action=a, draws=ref.step_draws(), state=steps[-1]["state"] if steps else init_state,
                                  obs=None, reward=-100.0, term=1, exc=1, info=None))
                break
            if t == 0 and want_boss is not None and int(state["boss_type"]) != want_boss:
                retry = True       # the boss is the reference's own tapped random.choice: take another seed
                break
            steps.append(dict(action=a, draws=ref.step_draws(), state=state,
                              obs=RefEnv.obs_record(obs), reward=float(r), term=int(term), exc=0, info=info))
            if term:
                break
        if retry:
            attempt += 1
            continue
        eps.append(dict(seed=seed, deck=deck, init_state=init_state, steps=steps))
        ep += 1
    T = max(len(e["steps"]) for e in eps)
    E = len(eps)
    out = dict(
        seeds=np.array([e["seed"] for e in eps], dtype=np.int64),
        decks=np.stack([e["deck"] for e in eps]).astype(np.uint8),
        init_state=np.zeros(E, dtype=L.STATE_DTYPE),
        length=np.array([len(e["steps"]) for e in eps], dtype=np.int32),
        action=np.full((T, E), -1, dtype=np.int32),
        draws=np.zeros((T, E), dtype=L.DRAWS_DTYPE),
        state=np.zeros((T, E), dtype=L.STATE_DTYPE),
        obs=np.zeros((T, E), dtype=L.OBS_DTYPE),
        reward=np.zeros((T, E), dtype=np.float64),
        term=np.zeros((T, E), dtype=np.uint8),
        exc=np.zeros((T, E), dtype=np.uint8),
        # info numerics: final_score, hand_type, chips, mult, has_error, played
        info=np.zeros((T, E, 6), dtype=np.int64),
    )
    for j, e in enumerate(eps):
        out["init_state"][j] = e["init_state"]
        for t, s in enumerate(e["steps"]):
            out["action"][t, j] = s["action"]
            out["draws"][t, j] = s["draws"]
            out["state"][t, j] = s["state"]
            if s["obs"] is not None:
                out["obs"][t, j] = s["obs"]
            out["reward"][t, j] = s["reward"]
            out["term"][t, j] = s["term"]
            out["exc"][t, j] = s["exc"]
            inf = s["info"] or {}
            if "final_score" in inf:
                out["info"][t, j] = [int(inf["final_score"]), int(inf["hand_type"]),
                                     int(inf["score_breakdown"]["final_chips"]), int(inf["score_breakdown"]["final_mult"]),
                                     0, 1]
            if "error" in inf:
                out["info"][t, j, 4] = 1
    return out


def main():
    cfgs = [("c1", 128, 101, 400), ("c3", 160, 20001, 500), ("c4", 160, 30001, 500), ("c4x", 0, 40001, 48)]
    only = sys.argv[1:]
    for name, episodes, seed0, max_steps in cfgs:
        if only and name not in only:
            continue
        out = record(name, episodes, seed0, max_steps)
        path = os.path.join(HERE, f"trace_{name}.npz")
        np.savez_compressed(path, **{k: (v.view(np.uint8) if v.dtype.fields else v) for k, v in out.items()})
        n = int(out["length"].sum())
        print(f"{name}: {out['action'].shape[1]} episodes, {n} reference steps, T={out['action'].shape[0]} -> {path} "
              f"({os.path.getsize(path)/1e6:.2f} MB)")


if __name__ == "__main__":
    main()
