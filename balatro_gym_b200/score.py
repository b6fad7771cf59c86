"""score_hands — batched hand scoring on the device (K1 + the joker interpreter K4).

Restates, per hand: BalatroGame._classify_hand (balatro_game.py:40-93),
CardAdapter.to_scoring_format chip values (balatro_env_2.py:287-325, cards.py:262-267) and
UnifiedScorer.score_hand with jokers given by id (unified_scoring.py:111-299,
complete_joker_effects.py:35-183).  All arguments are device tensors; no CPU fallback.
"""
from __future__ import annotations

from . import _lib


def score_hands(cards8, mods8=None, n_cards=None, jokers8=None, levels12=None, ctx=None, seed: int = 0,
                table_names: bool = False, want_x_mult: bool = True, want_money: bool = True, out=None, rules: bool = False):
    """cards8: uint8 [N,8] card codes; returns dict(hand_type u8, chips i32, mult i32, x_mult f64,
    score i64, money i32).  `out` may hold preallocated tensors with those keys.
    rules=True classifies with the rules evaluator (BalatroSimulator.evaluate_hand, balatro_sim.py:220-400:
    Five of a Kind / Flush House / Flush Five, Four Fingers and Shortcut read from jokers8) instead of the env's
    BalatroGame._classify_hand (BGYM_SCORE_RULES)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    assert cards8.dtype == torch.uint8 and cards8.dim() == 2 and cards8.shape[1] == 8 and cards8.is_contiguous()
    n, dev = cards8.shape[0], cards8.device
    if out is None:
        out = {"hand_type": torch.empty(n, dtype=torch.uint8, device=dev),
               "chips": torch.empty(n, dtype=torch.int32, device=dev),
               "mult": torch.empty(n, dtype=torch.int32, device=dev),
               "score": torch.empty(n, dtype=torch.int64, device=dev)}
        if want_x_mult:
            out["x_mult"] = torch.empty(n, dtype=torch.float64, device=dev)
        if want_money:
            out["money"] = torch.empty(n, dtype=torch.int32, device=dev)

    def p(t):
        return None if t is None else t.data_ptr()

    for t in (mods8, n_cards, jokers8, levels12, ctx):
        assert t is None or (t.is_contiguous() and t.device == dev)
    with torch.cuda.device(dev):
        rc = lib.bgym_score_hands(p(cards8), p(mods8), p(n_cards), p(jokers8), p(levels12), p(ctx),
                                  p(out["hand_type"]), p(out["chips"]), p(out["mult"]), p(out.get("x_mult")),
                                  p(out["score"]), p(out.get("money")), seed & 0xFFFFFFFF, n,
                                  (1 if table_names else 0) | (2 if rules else 0), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "bgym_score_hands")
    return out
