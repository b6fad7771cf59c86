"""Live differential tests of the C oracle against the unmodified reference (needs /root/reference
or oracle/_ref; skipped otherwise).  The committed golden traces cover the same ground offline."""
import numpy as np
import pytest

from oracle import coracle
from balatro_gym_b200 import layout as L


def test_lockstep_env(reference):
    import argparse
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import lockstep
    for cfg, seed0 in (("c1", 11), ("c3", 2222), ("c4", 3333)):
        args = argparse.Namespace(episodes=12, config=cfg, seed0=seed0, max_steps=1500, max_mismatch=0, invalid=True)
        steps, bad = lockstep.run(args)
        assert bad == 0 and steps > 300


def _random_hand(rng, n=None, with_mods=True):
    n = n or int(rng.integers(1, 9))
    codes = rng.choice(52, size=n, replace=False)
    mods = []
    for _ in range(n):
        enh = int(rng.integers(1, 9)) if (with_mods and rng.random() < 0.3) else 0
        ed = int(rng.integers(1, 4)) if (with_mods and rng.random() < 0.15) else 0
        seal = int(rng.integers(1, 5)) if (with_mods and rng.random() < 0.1) else 0
        mods.append((enh, ed, seal))
    return codes, mods


def _pack(hands):
    n = len(hands)
    cards = np.zeros((n, 8), np.uint8); mods = np.zeros((n, 8), np.uint16); nc = np.zeros(n, np.uint8)
    jk = np.zeros((n, 8), np.uint8); lv = np.ones((n, 12), np.uint8); ctx = np.zeros(n, L.SCORE_CTX_DTYPE)
    for i, h in enumerate(hands):
        k = len(h["codes"])
        nc[i] = k
        cards[i, :k] = h["codes"]
        mods[i, :k] = [e | (d << 4) | (s << 8) for e, d, s in h["mods"]]
        jk[i, :len(h["jokers"])] = h["jokers"]
        lv[i] = h["levels"]
        ctx[i]["hands_left"] = h["hands_left"]; ctx[i]["discards_left"] = h["discards_left"]
        ctx[i]["deck_len"] = h["deck_len"]; ctx[i]["use_replay"] = 1
        ctx[i]["bloodstone_bits"] = h["ref"]["bloodstone_bits"]
        if h["ref"]["misprint"]:
            ctx[i]["misprint"][0] = h["ref"]["misprint"][0]
    return cards, mods, nc, jk, lv, ctx


def make_score_cases(reference, n_cases, seed, table_names):
    from oracle.refscore import RefScorer
    rs = RefScorer()
    rng = np.random.default_rng(seed)
    hands = []
    for i in range(n_cases):
        codes, mods = _random_hand(rng)
        nj = int(rng.integers(0, 6))
        jokers = [int(x) for x in rng.choice(np.arange(1, 151), size=nj, replace=False)]
        levels = rng.integers(1, 6, size=12)
        h = dict(codes=codes, mods=mods, jokers=jokers, levels=levels, hands_left=int(rng.integers(1, 5)),
                 discards_left=int(rng.integers(0, 4)), deck_len=int(rng.integers(40, 53)))
        h["ref"] = rs.score(codes, mods, jokers, levels, h["hands_left"], h["discards_left"], h["deck_len"], table_names)
        hands.append(h)
    return hands


def check_scores(hands, out):
    for i, h in enumerate(hands):
        r = h["ref"]
        got = (int(out["hand_type"][i]), int(out["chips"][i]), int(out["mult"][i]), float(out["x_mult"][i]),
               int(out["score"][i]), int(out["money"][i]))
        exp = (r["hand_type"], r["chips"], r["mult"], r["x_mult"], r["score"], r["money"])
        assert got == exp, (i, h["codes"], h["mods"], h["jokers"], got, exp)


@pytest.mark.parametrize("table_names", [False, True])
def test_score_hands_vs_reference(reference, table_names):
    hands = make_score_cases(reference, 1500, 42 + table_names, table_names)
    cards, mods, nc, jk, lv, ctx = _pack(hands)
    out = coracle.score_hands(cards, mods, nc, jk, lv, ctx, flags=1 if table_names else 0)
    check_scores(hands, out)


def test_every_joker_row_vs_reference(reference):
    """One joker at a time against the reference, on hands built to trigger each family."""
    from oracle.refscore import RefScorer
    rs = RefScorer()
    rng = np.random.default_rng(7)
    hands = []
    special = [[0, 1, 2, 3], [51, 47, 43, 39, 35], [40, 41, 44, 45, 48], [8, 9, 10, 12, 16], [3, 7, 11, 15, 19],
               [44, 45, 46, 40, 41], [24, 25, 26, 27, 0], [0, 4, 8, 12, 16]]
    for jid in range(1, 151):
        for rep in range(6):
            if rep < len(special) and rep < 4:
                codes = np.array(special[(jid + rep) % len(special)])
                mods = [(0, 0, 0)] * len(codes)
            else:
                codes, mods = _random_hand(rng, with_mods=rep == 5)
            for tn in (False, True):
                h = dict(codes=codes, mods=mods, jokers=[jid], levels=np.ones(12, int), hands_left=1 + rep % 4,
                         discards_left=rep % 4, deck_len=52, tn=tn)
                h["ref"] = rs.score(codes, mods, [jid], None, h["hands_left"], h["discards_left"], 52, tn)
                hands.append(h)
    for tn in (False, True):
        sub = [h for h in hands if h["tn"] == tn]
        cards, mods, nc, jk, lv, ctx = _pack(sub)
        out = coracle.score_hands(cards, mods, nc, jk, lv, ctx, flags=1 if tn else 0)
        check_scores(sub, out)


def test_validator_port_accepts_the_reference(reference):
    """balatro_gym_b200.validate.BalatroEnvValidator is duck-typed: the unmodified reference passes it,
    so a failure on the device env is a real divergence and not a stricter check."""
    from balatro_gym_b200.validate import BalatroEnvValidator
    assert BalatroEnvValidator.validate_determinism(reference.BalatroEnv, seed=42, steps=100)
    assert BalatroEnvValidator.validate_action_masking(reference.BalatroEnv(seed=42))
