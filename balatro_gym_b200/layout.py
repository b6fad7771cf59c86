"""Record layouts of the C-ABI (include/bgym.h) as numpy structured dtypes, plus the
reference's enums/tables the host side needs.

Every dtype here mirrors one struct in include/bgym.h byte for byte (offsets are explicit and
checked against the library at load time by `_lib.check_layout`).
"""
from __future__ import annotations

import numpy as np

STATE_BYTES = 320
HOT_BYTES = 144
COLD_BYTES = 176
OBS_BYTES = 176
TOG_BYTES = 32      # BgymTog: device-only toggle record (authoritative copy of hot bytes 16..31 + select-path summary)
SEL_BYTES = 16      # BgymSel: device-only selection record (authoritative selected_cards + action_mask_bits)
OBS_DELTA_BYTES = 160   # what an observation delta carries of a record: all but the mask word (selection record) and padding
MIRROR_CORE_BYTES, MIRROR_SHOP_BYTES = 128, 32   # host mirror of the observations (bgym_scatter_dirty_obs)
MIRROR_CORE_CHUNKS, MIRROR_SHOP_CHUNKS = (0, 1, 2, 3, 4, 5, 8, 9), (6, 7)   # 16-byte chunks of an obs record
SYNC_TO_RECORDS, SYNC_FROM_RECORDS = 0, 1
INFO_BYTES = 32
DRAWS_BYTES = 256
NUM_ACTIONS = 60

# constants.py:34-39
PHASE_PLAY, PHASE_SHOP, PHASE_BLIND_SELECT, PHASE_PACK_OPEN = 0, 1, 2, 3
# constants.py:43-88
A_PLAY_HAND, A_DISCARD, A_SELECT_BASE, A_USE_CONS_BASE = 0, 1, 2, 10
A_SHOP_BUY_BASE, A_SHOP_REROLL, A_SHOP_END, A_SELL_JOKER_BASE = 20, 30, 31, 32
A_SELECT_BLIND_BASE, A_SKIP_BLIND = 45, 48

ERR_NONE, ERR_INVALID_ACTION, ERR_BOSS_RESTRICTION, ERR_CONSUMABLE_FAILED = 0, 1, 2, 3
ERR_SHOP, ERR_REF_EXCEPTION, ERR_UNSUPPORTED = 4, 5, 6
F_BEAT_BLIND, F_FAILED, F_GUARD_TERMINATED, F_PLAYED, F_AUTORESET_DONE, F_SHOP_DONE = 1, 2, 4, 8, 16, 32
FLAG_AUTORESET, FLAG_NO_OBS, FLAG_RANDOM_POLICY, FLAG_GEN_C3, FLAG_GEN_CONS = 1, 2, 4, 8, 16
# state generator of BASELINE configs[2] / configs[3] (include/bgym.h): applied by reset AND by the in-kernel autoreset
GENERATORS = {None: 0, "c3": FLAG_GEN_C3, "c4": FLAG_GEN_C3 | FLAG_GEN_CONS}
SCORE_TABLE_NAMES, SCORE_RULES = 1, 2


def _dt(fields, size):
    names, formats, offsets = zip(*fields)
    return np.dtype({"names": list(names), "formats": list(formats), "offsets": list(offsets),
                     "itemsize": size})


_HOT_FIELDS = [
    ("hand", "(8,)u1", 0), ("hand_code", "(8,)u1", 8),
    ("hand_n", "u1", 16), ("hand_size", "u1", 17), ("sel_n", "u1", 18), ("highlight_mask", "u1", 19),
    ("sel_order", "<u4", 20),
    ("face_down_mask", "u1", 24), ("phase", "u1", 25), ("round", "u1", 26), ("boss_type", "u1", 27),
    ("ep_len", "<u4", 28),
    ("joker_slots", "u1", 32), ("cons_slots", "u1", 33), ("n_magic_trick", "u1", 34), ("n_minimalist", "u1", 35),
    ("ante", "<i2", 36), ("jokers_sold", "<i2", 38),
    ("money", "<i4", 40), ("chips_needed", "<i4", 44),
    ("round_chips", "<i8", 48), ("chips_scored", "<i8", 56),
    ("best_hand", "<i4", 64), ("hands_played_total", "<i4", 68),
    ("hands_played_ante", "<i2", 72), ("boss_flags", "u1", 74), ("boss_cards_required", "u1", 75),
    ("boss_played_types", "<u2", 76), ("boss_hands_played", "u1", 78), ("deck_n", "u1", 79),
    ("boss_played_cards", "<u8", 80),
    ("joker_id", "(8,)u1", 88), ("cons_id", "(8,)u1", 96), ("hand_level", "(12,)u1", 104),
    ("shop_reroll_state", "<i4", 116), ("rng_seed", "<u4", 120), ("rng_ctr", "<u4", 124),
    ("hands_left", "u1", 128), ("discards_left", "u1", 129), ("joker_n", "u1", 130), ("cons_n", "u1", 131),
    ("episode", "<u4", 132), ("deck_extra", "(4,)<u2", 136),
]
_COLD_FIELDS = [
    ("deck", "(52,)<u2", 0), ("hand_play_count", "(12,)u1", 104),
    ("item_type", "(9,)u1", 116), ("item_id", "(9,)u1", 125), ("n_items", "u1", 134), ("deck_extra_n", "u1", 135),
    ("item_cost", "(9,)<i4", 136), ("reroll_cost", "<i4", 172),
]
HOT_DTYPE = _dt(_HOT_FIELDS, HOT_BYTES)
COLD_DTYPE = _dt(_COLD_FIELDS, COLD_BYTES)
# host-side combined record {hot, cold} back to back (checkpoints and tests)
STATE_DTYPE = _dt(_HOT_FIELDS + [(n, f, o + HOT_BYTES) for n, f, o in _COLD_FIELDS], STATE_BYTES)
HOT_FIELD_NAMES = [f[0] for f in _HOT_FIELDS]
COLD_FIELD_NAMES = [f[0] for f in _COLD_FIELDS]

OBS_DTYPE = _dt([
    ("hand", "(8,)i1", 0), ("selected_cards", "(8,)i1", 8), ("face_down_cards", "(8,)i1", 16),
    ("chips_scored", "<i8", 24), ("round_chips_scored", "<i4", 32), ("progress_ratio", "<f4", 36),
    ("mult", "<i4", 40), ("chips_needed", "<i4", 44), ("money", "<i4", 48), ("hands_played", "<i4", 52),
    ("best_hand_this_ante", "<i4", 56), ("ante", "<i2", 60), ("shop_rerolls", "<i2", 62),
    ("joker_ids", "(10,)<i2", 64), ("consumables", "(5,)<i2", 84), ("shop_items", "(10,)<i2", 94),
    ("shop_costs", "(10,)<i2", 114), ("hand_levels", "(12,)i1", 134),
    ("hand_size", "i1", 146), ("deck_size", "i1", 147), ("round", "i1", 148), ("hands_left", "i1", 149),
    ("discards_left", "i1", 150), ("joker_count", "i1", 151), ("joker_slots", "i1", 152),
    ("consumable_count", "i1", 153), ("consumable_slots", "i1", 154), ("phase", "i1", 155),
    ("boss_blind_active", "i1", 156), ("boss_blind_type", "i1", 157),
    ("action_mask_bits", "<u8", 160),
], OBS_BYTES)


TOG_DTYPE = _dt([
    ("hand_n", "u1", 0), ("hand_size", "u1", 1), ("sel_n", "u1", 2), ("highlight_mask", "u1", 3), ("sel_order", "<u4", 4),
    ("face_down_mask", "u1", 8), ("phase", "u1", 9), ("round", "u1", 10), ("boss_type", "u1", 11), ("ep_len", "<u4", 12),
    ("discards_left", "u1", 16), ("cons_n", "u1", 17), ("guard", "u1", 18), ("rng_seed", "<u4", 20),
], TOG_BYTES)
SEL_DTYPE = _dt([("selected_cards", "(8,)i1", 0), ("action_mask_bits", "<u8", 8)], SEL_BYTES)
# hot-record fields whose authoritative copy lives in the toggle record on the device
TOG_OWNED_FIELDS = ["hand_n", "hand_size", "sel_n", "highlight_mask", "sel_order", "face_down_mask", "phase", "round",
                    "boss_type", "ep_len"]


# host mirror core record (HostMirror.core): the observation fields that lie inside chunks 0..5, 8, 9, at their mirror offsets

def _mirror_core_dtype():
    fields = []
    for name in OBS_DTYPE.names:
        dt, off = OBS_DTYPE.fields[name][:2]
        lo, hi = off // 16, (off + dt.itemsize - 1) // 16
        if name in ("selected_cards", "action_mask_bits"):
            continue                                     # the selection record carries them
        if all(c in MIRROR_CORE_CHUNKS for c in range(lo, hi + 1)) and MIRROR_CORE_CHUNKS.index(hi) - MIRROR_CORE_CHUNKS.index(lo) == hi - lo:
            fields.append((name, dt, MIRROR_CORE_CHUNKS.index(lo) * 16 + off % 16))
    return np.dtype({"names": [f[0] for f in fields], "formats": [f[1] for f in fields], "offsets": [f[2] for f in fields],
                     "itemsize": MIRROR_CORE_BYTES})


MIRROR_CORE_DTYPE = _mirror_core_dtype()


def mask_from_bits(bits) -> np.ndarray:
    """The reference's obs['action_mask'] (int8[..., 60], balatro_env_2.py:1522) from the packed word(s)."""
    b = np.asarray(bits, dtype=np.uint64)
    return ((b[..., None] >> np.arange(NUM_ACTIONS, dtype=np.uint64)) & np.uint64(1)).astype(np.int8)


def obs_value(rec, key):
    """Field `key` of an observation record (array) in the reference's dict form; 'action_mask' is expanded."""
    return mask_from_bits(rec["action_mask_bits"]) if key == "action_mask" else rec[key]

# the 31 observation keys the reference emits, in its dict order (balatro_env_2.py:1488-1531)
OBS_KEYS = [
    "hand", "hand_size", "deck_size", "selected_cards", "chips_scored", "round_chips_scored",
    "progress_ratio", "mult", "chips_needed", "money", "ante", "round", "hands_left", "discards_left",
    "joker_count", "joker_ids", "joker_slots", "consumable_count", "consumables", "consumable_slots",
    "shop_items", "shop_costs", "shop_rerolls", "hand_levels", "phase", "action_mask", "hands_played",
    "best_hand_this_ante", "boss_blind_active", "boss_blind_type", "face_down_cards",
]

INFO_DTYPE = _dt([
    ("final_score", "<i8", 0), ("x_mult", "<f8", 8), ("chips", "<i4", 16), ("mult", "<i4", 20),
    ("hand_type", "i1", 24), ("error_code", "u1", 25), ("flags", "u1", 26), ("cards_played", "u1", 27),
    ("base_score", "<i4", 28),
], INFO_BYTES)

DRAWS_DTYPE = _dt([
    ("u", "(24,)<f8", 0), ("k", "(32,)u1", 192), ("n_u", "u1", 224), ("n_k", "u1", 225),
], DRAWS_BYTES)

SCORE_CTX_DTYPE = _dt([
    ("hands_left", "u1", 0), ("discards_left", "u1", 1), ("deck_len", "u1", 2),
    ("misprint", "(5,)u1", 3), ("bloodstone_bits", "u1", 8), ("use_replay", "u1", 9),
], 16)


# ---- card16 helpers ---------------------------------------------------------------------------
def card16(code, enhancement=0, edition=0, seal=0):
    """code(6b) | enhancement<<6 | edition<<10 | seal<<13 (include/bgym.h)."""
    return (code & 63) | ((enhancement & 15) << 6) | ((edition & 7) << 10) | ((seal & 7) << 13)


def card16_fields(c):
    return c & 63, (c >> 6) & 15, (c >> 10) & 7, (c >> 13) & 7


# ---- consumable names <-> ids (balatro_env_2.py:1545-1567) -------------------------------------
TAROT_NAMES = ['The Fool', 'The Magician', 'The High Priestess', 'The Empress', 'The Emperor',
               'The Hierophant', 'The Lovers', 'The Chariot', 'Strength', 'The Hermit',
               'Wheel of Fortune', 'Justice', 'The Hanged Man', 'Death', 'Temperance', 'The Devil',
               'The Tower', 'The Star', 'The Moon', 'The Sun', 'Judgement', 'The World']
PLANET_NAMES = ['Mercury', 'Venus', 'Earth', 'Mars', 'Jupiter', 'Saturn', 'Uranus', 'Neptune',
                'Pluto', 'Planet X', 'Ceres', 'Eris']
SPECTRAL_NAMES = ['Familiar', 'Grim', 'Incantation', 'Talisman', 'Aura', 'Wraith', 'Sigil', 'Ouija',
                  'Ectoplasm', 'Immolate', 'Ankh', 'Deja Vu', 'Hex', 'Trance', 'Medium', 'Cryptid',
                  'The Soul', 'Black Hole']


def consumable_id(name: str) -> int:
    if name in TAROT_NAMES:
        return 1 + TAROT_NAMES.index(name)
    if name in PLANET_NAMES:
        return 30 + PLANET_NAMES.index(name)
    if name in SPECTRAL_NAMES:
        return 50 + SPECTRAL_NAMES.index(name)
    # enum-style tarot names created by The Emperor (consumables.py:172)
    enum_style = [t.upper().replace(' ', '_') for t in TAROT_NAMES]
    if name in enum_style:
        return 100 + 1 + enum_style.index(name)
    raise KeyError(name)


def consumable_name(cid: int) -> str:
    if 1 <= cid <= 22:
        return TAROT_NAMES[cid - 1]
    if 30 <= cid <= 41:
        return PLANET_NAMES[cid - 30]
    if 50 <= cid <= 67:
        return SPECTRAL_NAMES[cid - 50]
    if 101 <= cid <= 122:
        return TAROT_NAMES[cid - 101].upper().replace(' ', '_')
    raise KeyError(cid)
