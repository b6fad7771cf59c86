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
from oracle.refenv import RefEnv  # noqa: E402
from balatro_gym_b200 import layout as L  # noqa: E402
from lockstep import inject_c3  # noqa: E402


class _FakeOV:  # inject_c3 writes the same mods into an oracle-side record; here we only need the ref side
    def __init__(self):
        self.state = np.zeros(1, dtype=L.STATE_DTYPE)


def record(config, episodes, seed0, max_steps, invalid_rate=0.02):
    rng = np.random.default_rng(seed0)
    ref = RefEnv(seed=1)
    eps = []
    for ep in range(episodes):
        seed = seed0 + ep
        obs, _ = ref.reset(seed)
        deck = ref.deck_codes()
        if config in ("c3", "c4"):
            fake = _FakeOV()
            fake.state[0]["deck"][:] = deck
            inject_c3(ref, fake, rng)
        init_state = ref.extract_state()
        steps = []
        first = True
        for t in range(max_steps):
            legal = ref.legal_actions()
            if config == "c1" and first:
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
            except OverflowError:
                break  # numpy-2-only int16 overflow of the reference's obs (pinned numpy 1.26 wraps)
            except Exception:
                # reference raises: expected -100 / terminated / state unchanged
                steps.append(dict(action=a, draws=ref.step_draws(), state=steps[-1]["state"] if steps else init_state,
                                  obs=None, reward=-100.0, term=1, exc=1, info=None))
                break
            steps.append(dict(action=a, draws=ref.step_draws(), state=ref.extract_state(),
                              obs=RefEnv.obs_record(obs), reward=float(r), term=int(term), exc=0, info=info))
            if term:
                break
        eps.append(dict(seed=seed, deck=deck, init_state=init_state, steps=steps))
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
    cfgs = [("c1", 128, 101, 400), ("c3", 160, 20001, 500), ("c4", 160, 30001, 500)]
    for name, episodes, seed0, max_steps in cfgs:
        out = record(name, episodes, seed0, max_steps)
        path = os.path.join(HERE, f"trace_{name}.npz")
        np.savez_compressed(path, **{k: (v.view(np.uint8) if v.dtype.fields else v) for k, v in out.items()})
        n = int(out["length"].sum())
        print(f"{name}: {episodes} episodes, {n} reference steps, T={out['action'].shape[0]} -> {path} "
              f"({os.path.getsize(path)/1e6:.2f} MB)")


if __name__ == "__main__":
    main()
