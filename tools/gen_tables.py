#!/usr/bin/env python
"""Generate include/bgym_tables.h from the reference's own data tables.

Run in the build container (where /root/reference is mounted):  python tools/gen_tables.py
The output is committed; nothing at run time reads /root/reference.

Numbers come from importing the reference, never from hand typing:
  JOKER_LIBRARY ids / names / base costs      balatro_gym/jokers.py:11-162
  BASE_HAND_VALUES                            balatro_gym/scoring_engine.py:27-40
  BLIND_CHIPS                                 balatro_gym/balatro_env_2.py:55-64
  COST_TABLE / ANTE_COST_MULT                 balatro_gym/shop.py:27-37
  fp64 power tables evaluated by CPython so device values are bit-identical:
      1.15**k  (shop.py:106)   0.8**n (boss_blinds.py:441)   1.5**k (balatro_env_2.py:73,
      complete_joker_effects.py:121)
The joker EFFECT table (what each joker does) is code in the reference
(complete_joker_effects.py:35-183), so it is restated here by hand, keyed by joker NAME, and
resolved to ids through JOKER_LIBRARY; tests/test_oracle_vs_reference.py checks every row against the
reference one joker at a time.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle.refenv import load_reference  # noqa: E402

R = load_reference()


def ident(name):
    s = re.sub(r"[^A-Za-z0-9]+", "_", name.replace("é", "e").replace("'", "")).strip("_").upper()
    return s


# ---- joker effect rows ---------------------------------------------------------------------------
# kind: how the row is evaluated
K_NONE, K_MAIN_ALWAYS, K_MAIN_HANDNAME, K_MAIN_SUIT_ANY, K_MAIN_STATE, K_MAIN_SPECIAL, \
    K_IND_RANKSET, K_IND_FACE, K_IND_SUIT = range(9)
# K_MAIN_STATE args
S_HALF, S_ABSTRACT, S_ACROBAT, S_MYSTIC, S_BANNER, S_BLUE, S_MISPRINT = range(7)
# K_MAIN_SPECIAL args
SP_BLACKBOARD, SP_SEEING_DOUBLE, SP_FLOWER_POT, SP_BARON, SP_SHOOT_MOON = range(5)
# hand-name ids used by K_MAIN_HANDNAME (the names complete_joker_effects.py:64-80 compares with)
HN_PAIR, HN_THREE_OAK, HN_TWO_PAIR, HN_STRAIGHT, HN_FLUSH, HN_FOUR_OAK = range(6)
SUIT = {"Clubs": 0, "Diamonds": 1, "Hearts": 2, "Spades": 3}


def rankset(ranks):
    m = 0
    for r in ranks:
        m |= 1 << r
    return m


FACE_JQK = rankset([11, 12, 13])
# name -> (kind, arg, chips, mult, xmult, money)
FX = {
    # main, unconditional                                   complete_joker_effects.py:39-53
    "Joker": (K_MAIN_ALWAYS, 0, 0, 4, 1.0, 0),
    "Stuntman": (K_MAIN_ALWAYS, 0, 250, 0, 1.0, 0),
    "Gros Michel": (K_MAIN_ALWAYS, 0, 0, 15, 1.0, 0),
    "Cavendish": (K_MAIN_ALWAYS, 0, 0, 0, 3.0, 0),
    "Popcorn": (K_MAIN_ALWAYS, 0, 0, 20, 1.0, 0),
    "Ice Cream": (K_MAIN_ALWAYS, 0, 100, 0, 1.0, 0),
    # main, game-state dependent                            :42-50
    "Misprint": (K_MAIN_STATE, S_MISPRINT, 0, 0, 1.0, 0),
    "Half Joker": (K_MAIN_STATE, S_HALF, 0, 20, 1.0, 0),
    "Abstract Joker": (K_MAIN_STATE, S_ABSTRACT, 0, 3, 1.0, 0),
    "Acrobat": (K_MAIN_STATE, S_ACROBAT, 0, 0, 3.0, 0),
    "Mystic Summit": (K_MAIN_STATE, S_MYSTIC, 0, 15, 1.0, 0),
    "Banner": (K_MAIN_STATE, S_BANNER, 30, 0, 1.0, 0),
    "Blue Joker": (K_MAIN_STATE, S_BLUE, 2, 0, 1.0, 0),
    # main, any scoring card of suit                        :56-61, :86-90
    "Greedy Joker": (K_MAIN_SUIT_ANY, SUIT["Diamonds"], 0, 3, 1.0, 0),
    "Lusty Joker": (K_MAIN_SUIT_ANY, SUIT["Hearts"], 0, 3, 1.0, 0),
    "Wrathful Joker": (K_MAIN_SUIT_ANY, SUIT["Spades"], 0, 3, 1.0, 0),
    "Gluttonous Joker": (K_MAIN_SUIT_ANY, SUIT["Clubs"], 0, 3, 1.0, 0),
    # main, hand name equality                              :64-80, :93-96
    "Jolly Joker": (K_MAIN_HANDNAME, HN_PAIR, 0, 8, 1.0, 0),
    "Zany Joker": (K_MAIN_HANDNAME, HN_THREE_OAK, 0, 12, 1.0, 0),
    "Mad Joker": (K_MAIN_HANDNAME, HN_TWO_PAIR, 0, 10, 1.0, 0),
    "Crazy Joker": (K_MAIN_HANDNAME, HN_STRAIGHT, 0, 12, 1.0, 0),
    "Droll Joker": (K_MAIN_HANDNAME, HN_FLUSH, 0, 10, 1.0, 0),
    "Sly Joker": (K_MAIN_HANDNAME, HN_PAIR, 50, 0, 1.0, 0),
    "Wily Joker": (K_MAIN_HANDNAME, HN_THREE_OAK, 100, 0, 1.0, 0),
    "Clever Joker": (K_MAIN_HANDNAME, HN_TWO_PAIR, 80, 0, 1.0, 0),
    "Devious Joker": (K_MAIN_HANDNAME, HN_STRAIGHT, 100, 0, 1.0, 0),
    "Crafty Joker": (K_MAIN_HANDNAME, HN_FLUSH, 80, 0, 1.0, 0),
    "The Duo": (K_MAIN_HANDNAME, HN_PAIR, 0, 0, 2.0, 0),
    "The Trio": (K_MAIN_HANDNAME, HN_THREE_OAK, 0, 0, 3.0, 0),
    "The Family": (K_MAIN_HANDNAME, HN_FOUR_OAK, 0, 0, 4.0, 0),
    "The Order": (K_MAIN_HANDNAME, HN_STRAIGHT, 0, 0, 3.0, 0),
    "The Tribe": (K_MAIN_HANDNAME, HN_FLUSH, 0, 0, 2.0, 0),
    # main, special                                         :99-127
    "Blackboard": (K_MAIN_SPECIAL, SP_BLACKBOARD, 0, 0, 3.0, 0),
    "Seeing Double": (K_MAIN_SPECIAL, SP_SEEING_DOUBLE, 0, 0, 2.0, 0),
    "Flower Pot": (K_MAIN_SPECIAL, SP_FLOWER_POT, 0, 0, 3.0, 0),
    "Baron": (K_MAIN_SPECIAL, SP_BARON, 0, 0, 1.5, 0),
    "Shoot the Moon": (K_MAIN_SPECIAL, SP_SHOOT_MOON, 0, 13, 1.0, 0),
    # individual, rank set                                  :139-147
    "Fibonacci": (K_IND_RANKSET, rankset([2, 3, 5, 8, 14]), 0, 8, 1.0, 0),
    "Even Steven": (K_IND_RANKSET, rankset([2, 4, 6, 8, 10]), 0, 4, 1.0, 0),
    "Odd Todd": (K_IND_RANKSET, rankset([3, 5, 7, 9, 14]), 31, 0, 1.0, 0),
    "Scholar": (K_IND_RANKSET, rankset([14]), 20, 4, 1.0, 0),
    "Walkie Talkie": (K_IND_RANKSET, rankset([4, 10]), 10, 4, 1.0, 0),
    "Wee Joker": (K_IND_RANKSET, rankset([2]), 8, 0, 1.0, 0),
    # individual, face cards                                :150-154  (arg = rank set that fires)
    "Scary Face": (K_IND_FACE, FACE_JQK, 30, 0, 1.0, 0),
    "Smiley Face": (K_IND_FACE, FACE_JQK, 0, 5, 1.0, 0),
    "Triboulet": (K_IND_FACE, rankset([12, 13]), 0, 0, 2.0, 0),
    # individual, suit                                      :157-162
    "Arrowhead": (K_IND_SUIT, SUIT["Spades"], 50, 0, 1.0, 0),
    "Onyx Agate": (K_IND_SUIT, SUIT["Clubs"], 0, 7, 1.0, 0),
    "Rough Gem": (K_IND_SUIT, SUIT["Diamonds"], 0, 0, 1.0, 1),
    "Bloodstone": (K_IND_SUIT, SUIT["Hearts"] | 0x80, 0, 0, 2.0, 0),  # 0x80: gated by a 50% roll
}


def main():
    lib = R.jokers.JOKER_LIBRARY
    by_name = {j.name: j for j in lib}
    for name in FX:
        assert name in by_name, name
    out = []
    w = out.append
    w("/* GENERATED by tools/gen_tables.py from the reference's data tables — do not edit.")
    w(" * Sources: balatro_gym/jokers.py:11-162, scoring_engine.py:27-40, balatro_env_2.py:55-64,")
    w(" * shop.py:27-37; fp64 powers evaluated by CPython (hex literals are exact). */")
    w("#ifndef BGYM_TABLES_H")
    w("#define BGYM_TABLES_H")
    w("#include <stdint.h>")
    w("")
    w("#define BGYM_NUM_JOKERS %d" % len(lib))
    # shop-eligible jokers (base_cost > 0, shop.py:126) are exactly ids 1..K: the kernels rely on it
    elig = [j.id for j in lib if j.base_cost > 0]
    assert elig == list(range(1, len(elig) + 1)), "shop-eligible joker ids are not contiguous from 1"
    w("#define BGYM_NUM_SHOP_JOKERS %d" % len(elig))
    for j in lib:
        w("#define BGYM_J_%s %d" % (ident(j.name), j.id))
    w("")
    w("/* base_cost by joker id (index 0 unused); 0 = legendary, not sold (jokers.py:9) */")
    costs = [0] * (len(lib) + 1)
    for j in lib:
        costs[j.id] = j.base_cost
    w("#define BGYM_JOKER_COST_INIT { %s }" % ", ".join(map(str, costs)))
    w("")
    w("/* (chips, mult) by HandType, level 1 (scoring_engine.py:27-40) */")
    bhv = R.scoring.BASE_HAND_VALUES
    w("#define BGYM_BASE_CHIPS_INIT { %s }" % ", ".join(str(bhv[h][0]) for h in R.HandType))
    w("#define BGYM_BASE_MULT_INIT  { %s }" % ", ".join(str(bhv[h][1]) for h in R.HandType))
    w("")
    w("/* BLIND_CHIPS[ante 1..8][small,big,boss] (balatro_env_2.py:55-64) */")
    bc = R.env_mod.BLIND_CHIPS
    rows = ["{ %d, %d, %d }" % (bc[a]["small"], bc[a]["big"], bc[a]["boss"]) for a in range(1, 9)]
    w("#define BGYM_BLIND_CHIPS_INIT { %s }" % ", ".join(rows))
    w("")
    ct = R.shop.COST_TABLE
    w("/* pack kind order: Standard, Joker, Tarot, Planet, Spectral (shop.py:27-33) */")
    w("#define BGYM_PACK_COST_INIT { %d, %d, %d, %d, %d }" % (
        ct["Standard Pack"], ct["Joker Pack"], ct["Tarot Pack"], ct["Planet Pack"], ct["Spectral Pack"]))
    w("#define BGYM_VOUCHER_COST_INIT { %d, %d }" % (ct["Voucher: Magic Trick"], ct["Voucher: Minimalist"]))
    w("#define BGYM_CARD_COST 40      /* shop.py:139 */")
    w("#define BGYM_REROLL_BASE 50    /* shop.py:101 */")
    w("")
    acm = R.shop.ANTE_COST_MULT
    w("/* ANTE_COST_MULT ** k, k = ante-1 in 0..100 (shop.py:106) */")
    w("#define BGYM_POW_1_15_INIT { %s }" % ", ".join((acm ** k).hex() for k in range(0, 101)))
    w("/* 0.8 ** n, n debuffed cards 0..8 (boss_blinds.py:441) */")
    w("#define BGYM_POW_0_8_INIT { %s }" % ", ".join((0.8 ** k).hex() for k in range(0, 9)))
    w("/* 1.5 ** k, k 0..100: blind scaling past ante 8 (balatro_env_2.py:73) and Baron")
    w(" * (complete_joker_effects.py:121) */")
    w("#define BGYM_POW_1_5_INIT { %s }" % ", ".join((1.5 ** k).hex() for k in range(0, 101)))
    w("")
    w("/* joker effect rows (hand-restated from complete_joker_effects.py:35-183, see tools/gen_tables.py) */")
    for i, n in enumerate(["NONE", "MAIN_ALWAYS", "MAIN_HANDNAME", "MAIN_SUIT_ANY", "MAIN_STATE",
                           "MAIN_SPECIAL", "IND_RANKSET", "IND_FACE", "IND_SUIT"]):
        w("#define BGYM_FX_%s %d" % (n, i))
    for i, n in enumerate(["HALF", "ABSTRACT", "ACROBAT", "MYSTIC", "BANNER", "BLUE", "MISPRINT"]):
        w("#define BGYM_FXS_%s %d" % (n, i))
    for i, n in enumerate(["BLACKBOARD", "SEEING_DOUBLE", "FLOWER_POT", "BARON", "SHOOT_MOON"]):
        w("#define BGYM_FXSP_%s %d" % (n, i))
    for i, n in enumerate(["PAIR", "THREE_OAK", "TWO_PAIR", "STRAIGHT", "FLUSH", "FOUR_OAK"]):
        w("#define BGYM_HN_%s %d" % (n, i))
    w("typedef struct BgymJokerFx { uint8_t kind; uint8_t _pad; uint16_t arg; int16_t chips; int16_t mult;"
      " float xmult; int16_t money; int16_t _pad2; } BgymJokerFx;")
    rows = []
    fx_by_id = {by_name[n].id: v for n, v in FX.items()}
    for jid in range(len(lib) + 1):
        k, a, c, m, x, mo = fx_by_id.get(jid, (K_NONE, 0, 0, 0, 1.0, 0))
        rows.append("{%d,0,%d,%d,%d,%.1ff,%d,0}" % (k, a, c, m, x, mo))
    w("#define BGYM_JOKER_FX_INIT { \\\n  %s }" % ", \\\n  ".join(
        ", ".join(rows[i:i + 6]) for i in range(0, len(rows), 6)))
    w("")
    w("#endif /* BGYM_TABLES_H */")
    path = os.path.join(REPO, "include", "bgym_tables.h")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
