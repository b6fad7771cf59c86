"""Known-answer vectors that pin the scoring formula.

The reference's own tests (tests/chips_test.py) are stale, but 8 of its 10 legacy known answers
(tests + balatro_gym/balatro_trajectories.json score deltas) still reproduce on the current
classify + to_scoring_format + UnifiedScorer path.  They are checked here against
the C oracle (CPU) and, in test_gpu_parity.py, against the CUDA kernel."""
import numpy as np

from oracle import coracle

R = {"2": 2, "3": 3, "4": 4, "5": 5, "6": 6, "7": 7, "8": 8, "9": 9, "T": 10, "J": 11, "Q": 12, "K": 13, "A": 14}
S = {"c": 0, "d": 1, "h": 2, "s": 3}


def code(card):
    return (R[card[0]] - 2) * 4 + S[card[1]]


KNOWN = [
    (["2s", "3s", "4s", "5s", "6s"], 8, 960),      # straight flush 2-6 spades      chips_test.py
    (["Td", "Jd", "Qd", "Kd", "Ad"], 8, 1208),     # straight flush 10-A diamonds    chips_test.py
    (["Ac", "2c", "3c", "4c", "5c"], 8, 1000),     # wheel straight flush clubs      chips_test.py
    (["Ac", "2d", "3c", "4c", "5c"], 4, 220),      # wheel straight                  chips_test.py
    (["As"], 0, 16),                               # lone ace                        chips_test.py
    (["Js", "Ts"], 0, 25),                         # balatro_trajectories.json
    (["5s", "4s", "3s", "2s"], 0, 19),             # balatro_trajectories.json
    (["Kc", "Qc"], 0, 25),                         # balatro_trajectories.json
]


def known_arrays():
    cards = np.zeros((len(KNOWN), 8), dtype=np.uint8)
    n = np.zeros(len(KNOWN), dtype=np.uint8)
    for i, (cs, _, _) in enumerate(KNOWN):
        n[i] = len(cs)
        cards[i, :len(cs)] = [code(c) for c in cs]
    ht = np.array([k[1] for k in KNOWN])
    score = np.array([k[2] for k in KNOWN])
    return cards, n, ht, score


def test_known_answers_oracle():
    cards, n, ht, score = known_arrays()
    out = coracle.score_hands(cards, n_cards=n)
    assert out["hand_type"].tolist() == ht.tolist()
    assert out["score"].tolist() == score.tolist()


def test_no_five_of_a_kind_family():
    # current _classify_hand never returns FIVE_KIND / FLUSH_HOUSE / FLUSH_FIVE:
    # five aces of one suit classify as FLUSH with score (35 + 55) * 4 = 360
    cards = np.zeros((1, 8), dtype=np.uint8)
    cards[0, :5] = code("As")
    out = coracle.score_hands(cards)
    assert int(out["hand_type"][0]) == 5 and int(out["score"][0]) == 360


def test_classification_priorities():
    cases = [
        (["2c", "2d", "2h", "2s", "9c"], 7), (["2c", "2d", "2h", "9s", "9c"], 6), (["2c", "4c", "6c", "8c", "Tc"], 5),
        (["2c", "3d", "4h", "5s", "6c"], 4), (["2c", "2d", "2h", "8s", "9c"], 3), (["2c", "2d", "8h", "8s", "9c"], 2),
        (["2c", "2d", "7h", "8s", "9c"], 1), (["2c", "4d", "7h", "8s", "9c"], 0),
        # 8-card plays (no 5-card cap): two trips -> counts [3,3,2] is THREE_KIND, not FULL_HOUSE
        (["2c", "2d", "2h", "3c", "3d", "3h", "4c", "4d"], 3),
        # flush needs ONE suit among all cards; 6 cards with one off-suit is not a flush
        (["2c", "4c", "6c", "8c", "Tc", "Qd"], 0),
        # straight inside 7 cards
        (["2c", "3d", "4h", "5s", "6c", "9d", "Kd"], 4),
    ]
    cards = np.zeros((len(cases), 8), dtype=np.uint8)
    n = np.zeros(len(cases), dtype=np.uint8)
    for i, (cs, _) in enumerate(cases):
        n[i] = len(cs)
        cards[i, :len(cs)] = [code(c) for c in cs]
    out = coracle.score_hands(cards, n_cards=n)
    assert out["hand_type"].tolist() == [c[1] for c in cases]
