"""O2: drive the reference's scoring pipeline directly — TEST INFRASTRUCTURE.

classify (balatro_game.py:40-93) + CardAdapter.to_scoring_format (balatro_env_2.py:287-325) +
UnifiedScorer.score_hand (unified_scoring.py:111-299) with jokers given as NAME strings so that
CompleteJokerEffects actually fires (unified_scoring.py:164-165).  The module-global `random` of
complete_joker_effects is tapped so Misprint / Bloodstone draws can be replayed into the kernels.
"""
from __future__ import annotations

import numpy as np

from .refenv import load_reference, TapRandom

_TABLE_NAME = {1: "Pair", 3: "Three of a Kind", 7: "Four of a Kind"}


class RefScorer:
    def __init__(self, global_seed=999):
        self.R = R = load_reference()
        self.log = []
        self.rng = TapRandom(global_seed, self.log)
        R.jeff.random = self.rng
        self.engine = R.scoring.ScoreEngine()
        self.scorer = R.unified.UnifiedScorer(self.engine, R.jeff.CompleteJokerEffects())
        self.game = R.game.BalatroGame(self.engine)

    def score(self, codes, mods=None, joker_ids=(), levels=None, hands_left=4, discards_left=3, deck_len=52,
              table_names=False):
        """codes: list of card codes; mods: list of (enh, edition, seal) or None.
        Returns dict(hand_type, chips, mult, x_mult, score, money, misprint[list], bloodstone_bits)."""
        R = self.R
        C = R.cards
        self.R.jeff.random = self.rng
        n = len(codes)
        cards = [C.Card(rank=C.Rank(c // 4 + 2), suit=C.Suit(c % 4)) for c in codes]
        st = R.env_mod.UnifiedGameState()
        for i in range(n):
            enh, ed, seal = mods[i] if mods is not None else (0, 0, 0)
            if enh or ed or seal:
                st.card_states[i] = C.CardState(i, C.Enhancement(enh), C.Edition(ed), C.Seal(seal))
        scoring = [R.env_mod.CardAdapter.to_scoring_format(cards[i], i, st) for i in range(n)]
        for ht in R.HandType:
            self.engine.hand_levels[ht] = 1 if levels is None else min(15, max(1, int(levels[int(ht)])))
        hand_type, _ = self.game._classify_hand(cards)
        if table_names and int(hand_type) in _TABLE_NAME:
            name = _TABLE_NAME[int(hand_type)]
        else:
            name = hand_type.name.replace("_", " ").title()
        names = [R.JOKER_BY_ID[j].name for j in joker_ids if j]
        gs = {"jokers": names, "deck": [None] * deck_len, "hands_left": hands_left,
              "discards_left": discards_left, "money": 0}
        self.log.clear()
        ctx = R.unified.ScoringContext(cards=scoring, scoring_cards=scoring, hand_type=hand_type,
                                       hand_type_name=name, game_state=gs)
        score, bd = self.scorer.score_hand(ctx)
        # split the tapped draws: individual phase = one u per (card, joker) [+1 when 8 Ball meets an 8],
        # main phase = one randint(0,23) per joker
        us = [v for t, v in self.log if t == "u"]
        ks = [v for t, v in self.log if t == "k"]
        bits, iu = 0, 0
        for c in range(n):
            rank = scoring[c].rank
            for nm in names:
                roll = us[iu]; iu += 1
                if nm == "Bloodstone" and roll < 0.5:
                    bits |= 1 << c
                if nm == "8 Ball" and rank == 8:
                    iu += 1
        assert iu == len(us), (iu, len(us))
        misprint = [ks[j] for j, nm in enumerate(names) if nm == "Misprint"]
        return dict(hand_type=int(hand_type), chips=int(bd["final_chips"]), mult=int(bd["final_mult"]),
                    x_mult=float(bd["final_x_mult"]), score=int(score), money=int(bd["money_gained"]),
                    misprint=misprint, bloodstone_bits=bits)
