"""BGYM_SCORE_RULES (the pinnable half): the rules evaluator `BalatroSimulator.evaluate_hand`
(balatro_gym/balatro_sim.py:220-400, UNMODIFIED, imported from /root/reference or oracle/_ref) against the C oracle on
random 1-8-card multisets with repeated cards, with and without Four Fingers / Shortcut — and, on the GPU, the CUDA
kernel against both."""
import numpy as np
import pytest

from balatro_gym_b200 import layout as L
from oracle import coracle

_TOP = {"High Card": 0, "Pair": 1, "Two Pair": 2, "Three of a Kind": 3, "Straight": 4, "Flush": 5, "Full House": 6,
        "Four of a Kind": 7, "Straight Flush": 8, "Five of a Kind": 9, "Flush House": 10, "Flush Five": 11}
_SUITS = ["Clubs", "Diamonds", "Hearts", "Spades"]          # code & 3 (cards.py:103); the evaluator only compares them
J_FOUR_FINGERS, J_SHORTCUT = 18, 69


def random_multisets(n, seed):
    """1-8 cards; half of the hands are drawn from a small pool of ranks / suits so that repeated ranks, repeated
    cards, flushes and straights are common."""
    rng = np.random.default_rng(seed)
    nc = rng.integers(1, 9, n).astype(np.uint8)
    nc[: n // 2] = rng.integers(4, 7, n // 2)                  # where flushes / straights can exist
    rank = rng.integers(0, 13, (n, 8))
    suit = rng.integers(0, 4, (n, 8))
    few = rng.random(n) < 0.5
    base = rng.integers(0, 13, n)
    narrow = (base[:, None] + rng.integers(0, 6, (n, 8))) % 13                       # ranks within a window of 6
    rank = np.where(few[:, None], narrow, rank)
    dup = rng.random((n, 8)) < 0.35
    rank = np.where(dup, rank[:, :1], rank)                                            # copies of the first card's rank
    one_suit = rng.random(n) < 0.4
    suit = np.where(one_suit[:, None] & (rng.random((n, 8)) < 0.9), suit[:, :1], suit)
    cards = (rank * 4 + suit).astype(np.uint8)
    jk = np.zeros((n, 8), np.uint8)
    jk[:, 0] = np.where(rng.random(n) < 0.4, J_FOUR_FINGERS, 0)
    jk[:, 3] = np.where(rng.random(n) < 0.4, J_SHORTCUT, 0)
    jk[:, 5] = rng.integers(1, 151, n) * (rng.random(n) < 0.3)                        # an unrelated joker
    return cards, nc, jk


def reference_tops(cards, nc, jk):
    from oracle.refenv import load_rules_evaluator
    sim_mod = load_rules_evaluator()
    sim = sim_mod.BalatroSimulator()
    out = np.zeros(len(cards), np.uint8)
    for i in range(len(cards)):
        sim.player_state.jokers = [int(j) for j in jk[i] if j]
        hand = [sim_mod.Card(rank=int(c) // 4 + 2, suit=_SUITS[int(c) & 3]) for c in cards[i, :nc[i]]]
        out[i] = _TOP[sim.evaluate_hand(hand)["top"]]
    return out


def test_oracle_rules_classifier_matches_the_reference(reference):
    n = 120_000
    cards, nc, jk = random_multisets(n, 7)
    ref = reference_tops(cards, nc, jk)
    got = coracle.score_hands(cards, n_cards=nc, jokers8=jk, flags=L.SCORE_RULES)["hand_type"]
    bad = np.flatnonzero(ref != got)
    assert len(bad) == 0, (len(bad), cards[bad[0], :nc[bad[0]]].tolist(), jk[bad[0]].tolist(), int(ref[bad[0]]), int(got[bad[0]]))
    assert set(np.unique(ref)) == set(range(12))          # every hand type of the evaluator occurs, the three new ones too
    # the three obvious hands of VERDICT r01
    five = np.array([[48, 48, 48, 48, 48, 0, 0, 0], [48, 49, 48, 50, 51, 0, 0, 0], [48, 48, 48, 44, 44, 0, 0, 0]], np.uint8)
    tops = coracle.score_hands(five, flags=L.SCORE_RULES)["hand_type"].tolist()
    assert tops == [11, 9, 10] and reference_tops(five, np.full(3, 5), np.zeros((3, 8), np.uint8)).tolist() == tops


@pytest.mark.gpu
def test_cuda_rules_classifier(reference):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from balatro_gym_b200 import score_hands
    n = 1 << 18
    cards, nc, jk = random_multisets(n, 11)
    t = lambda a: torch.from_numpy(a).cuda()
    out = score_hands(t(cards), n_cards=t(nc), jokers8=t(jk), rules=True)
    exp = coracle.score_hands(cards, n_cards=nc, jokers8=jk, flags=L.SCORE_RULES)
    for k in ("hand_type", "chips", "mult", "x_mult", "score", "money"):
        assert np.array_equal(exp[k], out[k].cpu().numpy()), k
    m = 100_000                                             # and straight against the unmodified evaluator
    assert np.array_equal(reference_tops(cards[:m], nc[:m], jk[:m]), out["hand_type"][:m].cpu().numpy())
    assert set(np.unique(exp["hand_type"])) == set(range(12))
