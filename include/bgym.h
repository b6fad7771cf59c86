/* bgym.h — C-ABI of the B200-native batched Balatro environment step path.
 *
 * This is the drop-in boundary for the hot path of cassiusfive/balatro-gym:
 *   BalatroEnv.reset            balatro_gym/balatro_env_2.py:505-558   -> bgym_reset
 *   BalatroEnv.step             balatro_gym/balatro_env_2.py:616-1064  -> bgym_step
 *   BalatroEnv._get_action_mask balatro_gym/balatro_env_2.py:1426-1471 -> bgym_action_mask
 *   BalatroEnv._get_observation balatro_gym/balatro_env_2.py:1473-1541 -> BgymObs + BgymSel (written by reset/step)
 *   _classify_hand + CardAdapter.to_scoring_format + UnifiedScorer.score_hand
 *       balatro_gym/balatro_game.py:40-93, balatro_env_2.py:287-325,
 *       balatro_gym/unified_scoring.py:111-299                         -> bgym_score_hands
 *
 * The reference has no FFI of its own (it is pure Python behind the Gymnasium
 * Env protocol); INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - All pointers are plain borrowed pointers; the library never allocates or
 *     frees caller memory.  Entry points named *_host take HOST pointers and do
 *     their own staging through a BgymVec handle; all others take DEVICE pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     Device entry points are asynchronous and stream ordered.
 *   - Return value: 0 on success, >0 a cudaError_t, <0 an argument error
 *     (BGYM_E_*).  bgym_last_error() gives a thread-local message.
 *   - Records are fixed-layout little-endian structs (below); arrays of records
 *     are dense AoS: record i starts at base + i * sizeof(record).
 */
#ifndef BGYM_H
#define BGYM_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGYM_ABI_VERSION 2

/* ---- sizes ------------------------------------------------------------- */
#define BGYM_STATE_BYTES 320
#define BGYM_HOT_BYTES   144
#define BGYM_COLD_BYTES  176
#define BGYM_OBS_BYTES   176
#define BGYM_TOG_BYTES   32    /* BgymTog: the card-select working set of an env (device only) */
#define BGYM_SEL_BYTES   16    /* BgymSel: the two observation fields a card select changes (device only) */
#define BGYM_OBS_DELTA_BYTES 160   /* bytes 0..159 of a BgymObs: all of it but the mask word (carried by BgymSel) and padding */
#define BGYM_INFO_BYTES  32
#define BGYM_DRAWS_BYTES 256
#define BGYM_NUM_ACTIONS 60
#define BGYM_FEATURE_DIM 448   /* bgym_featurize row: 416 one-hot + 10 joker ids + 21 scalars + 1 pad */
#define BGYM_DT_F32  0
#define BGYM_DT_BF16 1
#define BGYM_NUM_HAND_TYPES 12
#define BGYM_MAX_HAND 8
#define BGYM_DECK_SLOTS 52

/* ---- enums (values follow the reference) -------------------------------- */
/* constants.py:34-39 */
enum { BGYM_PHASE_PLAY = 0, BGYM_PHASE_SHOP = 1, BGYM_PHASE_BLIND_SELECT = 2, BGYM_PHASE_PACK_OPEN = 3 };
/* constants.py:43-88 */
enum {
  BGYM_A_PLAY_HAND = 0, BGYM_A_DISCARD = 1, BGYM_A_SELECT_BASE = 2, BGYM_A_USE_CONS_BASE = 10,
  BGYM_A_SHOP_BUY_BASE = 20, BGYM_A_SHOP_REROLL = 30, BGYM_A_SHOP_END = 31,
  BGYM_A_SELL_JOKER_BASE = 32, BGYM_A_SELL_CONS_BASE = 37, BGYM_A_SELECT_BLIND_BASE = 45,
  BGYM_A_SKIP_BLIND = 48, BGYM_A_PACK_BASE = 50, BGYM_A_SKIP_PACK = 55
};
/* scoring_engine.py:12-24 */
enum {
  BGYM_HT_HIGH_CARD = 0, BGYM_HT_ONE_PAIR, BGYM_HT_TWO_PAIR, BGYM_HT_THREE_KIND, BGYM_HT_STRAIGHT,
  BGYM_HT_FLUSH, BGYM_HT_FULL_HOUSE, BGYM_HT_FOUR_KIND, BGYM_HT_STRAIGHT_FLUSH,
  BGYM_HT_FIVE_KIND, BGYM_HT_FLUSH_HOUSE, BGYM_HT_FLUSH_FIVE
};
/* cards.py:63-91 — card16 = code(6b) | enhancement<<6 (4b) | edition<<10 (3b) | seal<<13 (3b)
 * code = (rank-2)*4 + suit (cards.py:103), suit: clubs 0 diamonds 1 hearts 2 spades 3 */
enum { BGYM_ENH_NONE = 0, BGYM_ENH_BONUS, BGYM_ENH_MULT, BGYM_ENH_WILD, BGYM_ENH_GLASS, BGYM_ENH_STEEL,
       BGYM_ENH_STONE, BGYM_ENH_GOLD, BGYM_ENH_LUCKY };
enum { BGYM_ED_NONE = 0, BGYM_ED_FOIL, BGYM_ED_HOLO, BGYM_ED_POLY, BGYM_ED_NEGATIVE };
enum { BGYM_SEAL_NONE = 0, BGYM_SEAL_GOLD, BGYM_SEAL_RED, BGYM_SEAL_BLUE, BGYM_SEAL_PURPLE };
/* shop.py:17-21 */
enum { BGYM_ITEM_PACK = 1, BGYM_ITEM_CARD = 2, BGYM_ITEM_JOKER = 3, BGYM_ITEM_VOUCHER = 4 };
/* shop item_id for packs / vouchers (shop.py:27-35 key order) */
enum { BGYM_PACK_STANDARD = 0, BGYM_PACK_JOKER = 1, BGYM_PACK_TAROT = 2, BGYM_PACK_PLANET = 3, BGYM_PACK_SPECTRAL = 4 };
enum { BGYM_VOUCHER_MAGIC_TRICK = 0, BGYM_VOUCHER_MINIMALIST = 1 };

/* consumable ids = the reference's observation id map (balatro_env_2.py:1545-1567):
 *   tarots 1..22, planets 30..41, spectrals 50..67.  Tarots created by The Emperor carry
 *   enum-style names ('THE_FOOL', consumables.py:172) that resolve on use but observe as 0:
 *   they are stored as 100 + tarot id. */
#define BGYM_CONS_TAROT_BASE     1
#define BGYM_CONS_PLANET_BASE    30
#define BGYM_CONS_SPECTRAL_BASE  50
#define BGYM_CONS_ENUMSTYLE_BASE 100

/* error codes in BgymInfo.error_code */
enum {
  BGYM_ERR_NONE = 0,
  BGYM_ERR_INVALID_ACTION = 1,   /* masked action: reward -1.0, state unchanged (balatro_env_2.py:626-627) */
  BGYM_ERR_BOSS_RESTRICTION = 2, /* can_play_hand false: reward -1.0 (balatro_env_2.py:677-680) */
  BGYM_ERR_CONSUMABLE_FAILED = 3,/* use_consumable success False: reward -1.0 (balatro_env_2.py:1167-1169) */
  BGYM_ERR_SHOP = 4,             /* shop.step error (joker slots full, reroll unaffordable): reward -1.0 */
  BGYM_ERR_REF_EXCEPTION = 5,    /* the reference raises here; we return the
                                    SafeBalatroEnv convention: reward -100.0, terminated, state unchanged
                                    (train_balatro_fixed.py:262-269) */
  BGYM_ERR_UNSUPPORTED = 6       /* a Cryptid that would exceed the 4 appended cards deck_extra holds: reward -1.0,
                                    state unchanged (every consumable of the reference is otherwise implemented) */
};
/* BgymInfo.flags */
enum { BGYM_F_BEAT_BLIND = 1, BGYM_F_FAILED = 2, BGYM_F_GUARD_TERMINATED = 4, BGYM_F_PLAYED = 8,
       BGYM_F_AUTORESET_DONE = 16, BGYM_F_SHOP_DONE = 32 };

/* flags argument of bgym_reset / bgym_step */
enum {
  BGYM_FLAG_AUTORESET = 1,  /* step: a terminated env is re-initialised in place (native Philox shuffle)
                               and the returned observation is the first of the new episode */
  BGYM_FLAG_NO_OBS = 2,     /* skip observation emission (obs may be NULL) */
  BGYM_FLAG_RANDOM_POLICY = 4,/* step: every env draws its own uniform random LEGAL action from its mask
                               (the policy the reference is benchmarked with); `actions` becomes an
                               OUTPUT array that receives the chosen actions */
  /* Synthetic-state generator of BASELINE configs[2]/[3] (the reference-side
   * counterpart is the injection recipe of SURVEY Appendix E, oracle/refbaseline.py::_inject_c3).
   * Honoured by bgym_reset AND by the in-kernel autoreset of bgym_step, so every episode of a long
   * rollout starts from a generated state, not only the first.  All draws come from Philox4x32-10
   * keyed (episode seed, BGYM_GEN_KEY1); integer draws are floor(word * n / 2^bits) (|p - 1/n| < n / 2^27).
   *   card k of the freshly built deck (k = suit * 13 + rank - 2, balatro_env_2.py:519-522): block counter k / 2,
   *   words (w0, w1) = (x, y) for even k, (z, w) for odd k:
   *     enhancement: top two bits of w0 zero (p = 1/4) -> 1 + the next three bits (uniform over the 8)
   *     edition:     e = floor((w0 & 0x7FFFFFF) * 30 / 2^27) < 3 (p = 1/10) -> 1 + e (FOIL, HOLO, POLY)
   *     seal:        s = floor(w1 * 40 / 2^32) < 4 (p = 1/10) -> 1 + s
   *     the modifiers travel with the card through the shuffle (i.i.d., so position-keyed is the same law)
   *   jokers: blocks 64, 65: five draws without replacement over the BGYM_NUM_SHOP_JOKERS ids with
   *     base_cost > 0 (draw t picks the floor(w_t * (145 - t) / 2^32)-th id not chosen yet), in draw order */
  BGYM_FLAG_GEN_C3 = 8,
  /* with BGYM_FLAG_GEN_C3: additionally fill both consumable slots, each uniform over the 52 consumable
   * names of the reference (tarots 1..22, planets 30..41, spectrals 50..67): words y, z of block 65 */
  BGYM_FLAG_GEN_CONS = 16
};
#define BGYM_GEN_KEY1 0xB200C3C4u
/* flags argument of bgym_score_hands */
enum {
  BGYM_SCORE_TABLE_NAMES = 1,/* hand names as complete_joker_effects.py:64-80 expects ('Pair',
                                'Three of a Kind', 'Four of a Kind'); default = the env's own names
                                ('One Pair', 'Three Kind', 'Four Kind', balatro_env_2.py:674) */
  BGYM_SCORE_RULES = 2       /* classify with the RULES evaluator, BalatroSimulator.evaluate_hand
                                (balatro_sim.py:110-400) instead of BalatroGame._classify_hand: FIVE_KIND /
                                FLUSH_HOUSE / FLUSH_FIVE exist, a rank group counts only at EXACTLY 5/4/3/2 cards,
                                flushes and straights exist for at most 5 played cards, the Four Fingers joker
                                (4-card flushes and straights) and the Shortcut joker (one rank gap) are read from
                                jokers8, and cards8 may repeat a card.  hand_type = the evaluator's 'top' entry;
                                base chips / mult and the joker pipeline then run on that hand type. */
};

#define BGYM_E_ARG     (-1)
#define BGYM_E_NODEV   (-2)
#define BGYM_E_ALIGN   (-3)

/* ---- per-env state (320 B) = hot record (144 B) + cold record (176 B) ----------
 * Restates UnifiedGameState (balatro_env_2.py:166-211) + BalatroGame (balatro_game.py:16-28)
 * + ScoreEngine levels/counts (scoring_engine.py:65-69) + BossBlindManager.blind_state
 * (boss_blinds.py:311-319) + Shop inventory (shop.py:96-148).  SURVEY.md Appendix D.
 *
 * ON THE DEVICE the two halves live in two dense arrays, hot[n] (BgymHot, 144 B = 9 x 16) and
 * cold[n] (BgymCold, 176 B = 11 x 16), next to tog[n] (BgymTog, 32 B): card-select toggles (~75 % of steps) touch only
 * the toggle array, and both record strides are odd multiples of 16 B so a tile staged in shared memory is
 * bank-conflict free for 128-bit per-lane access.  BgymState = {hot, cold} back to back is the
 * HOST-side record (checkpoints, bgym_vec_get_state/set_state, the test oracle). */
/* hot record, 144 B (offset: field)
 *   0 hand[8]             hand_indexes: deck index per hand slot, 0xFF = empty
 *   8 hand_code[8]        cache: card code of deck[hand[i]], 0xFF = none (obs['hand'])
 *  16 hand_n, 17 hand_size, 18 sel_n, 19 highlight_mask (game.highlighted_indexes as a slot bit set)
 *  20 sel_order           selected_cards, ordered: nibble k = slot of the k-th selection
 *  24 face_down_mask, 25 phase, 26 round (1 small 2 big 3 boss), 27 boss_type (BossBlindType, 0 none)
 *  28 ep_len              valid steps this episode.  Bytes 16..31 are everything a card toggle changes; ON THE DEVICE
 *                         their authoritative copy is the env's toggle record (BgymTog below): a toggle updates it
 *                         there only, and bgym_sync_state folds it back into the hot record
 *  32 joker_slots, 33 cons_slots, 34 n_magic_trick, 35 n_minimalist (voucher counts)
 *  36 ante i16, 38 jokers_sold i16, 40 money i32, 44 chips_needed i32
 *  48 round_chips i64 (round_chips_scored), 56 chips_scored i64
 *  64 best_hand i32 (best_hand_this_ante, saturating), 68 hands_played_total i32
 *  72 hands_played_ante i16, 74 boss_flags (bit0 = blind_state['first_hand']),
 *  75 boss_cards_required (Verdant), 76 boss_played_types u16 (bit set over HandType),
 *  78 boss_hands_played, 79 deck_n (len(deck)), 80 boss_played_cards u64 (bit set over deck indices)
 *  88 joker_id[8] (JOKER_LIBRARY ids, 0 empty), 96 cons_id[8] (consumable ids, 0 empty)
 * 104 hand_level[12]      state.hand_levels (uncapped); engine level = min(level, 15)
 * 116 shop_reroll_state   state.shop_reroll_cost (stale copy used by the mask)
 * 120 rng_seed, 124 rng_ctr (native Philox key word / block counter)
 * 128 hands_left, 129 discards_left, 130 joker_n, 131 cons_n
 * 132 episode (in-kernel autoresets so far)
 * 136 deck_extra[4] u16   cards appended to the deck list by Cryptid (consumables.py:582-592), oldest first:
 *                         card code | the target's modifiers at copy time (card16 layout; the modifiers only serve the
 *                         dataclass equality list.remove uses in Immolate).  The deck is [surviving original cards] ++
 *                         [deck_extra[0..deck_extra_n)]; deck_n = len(deck) counts both.  Capacity: 4 appended cards
 *                         (two Cryptid uses outstanding); a Cryptid beyond that returns BGYM_ERR_UNSUPPORTED. */
#define BGYM_HOT_FIELDS \
  uint8_t hand[8]; uint8_t hand_code[8]; \
  uint8_t hand_n; uint8_t hand_size; uint8_t sel_n; uint8_t highlight_mask; uint32_t sel_order; \
  uint8_t face_down_mask; uint8_t phase; uint8_t round; uint8_t boss_type; \
  uint32_t ep_len; \
  uint8_t joker_slots; uint8_t cons_slots; uint8_t n_magic_trick; uint8_t n_minimalist; \
  int16_t ante; int16_t jokers_sold; int32_t money; int32_t chips_needed; \
  int64_t round_chips; int64_t chips_scored; int32_t best_hand; int32_t hands_played_total; \
  int16_t hands_played_ante; uint8_t boss_flags; uint8_t boss_cards_required; \
  uint16_t boss_played_types; uint8_t boss_hands_played; uint8_t deck_n; uint64_t boss_played_cards; \
  uint8_t joker_id[8]; uint8_t cons_id[8]; uint8_t hand_level[12]; int32_t shop_reroll_state; \
  uint32_t rng_seed; uint32_t rng_ctr; uint8_t hands_left; uint8_t discards_left; uint8_t joker_n; uint8_t cons_n; \
  uint32_t episode; uint16_t deck_extra[4];

/* cold record, 176 B (offset: field)
 *   0 deck[52] u16        card16 per deck index: code of the card now at that index of the deck list (0 beyond
 *                         deck_n) | the modifiers of state.card_states[index] (they are keyed by INDEX in the
 *                         reference, so Immolate moves codes but not modifiers)
 * 104 hand_play_count[12] engine.hand_play_counts, saturating at 255 (never read by the reference's
 *                         step path; kept for save_state)
 * 116 item_type[9], 125 item_id[9] (joker id / pack kind / voucher kind / card int), 134 n_items,
 * 135 deck_extra_n (valid entries of the hot record's deck_extra)
 * 136 item_cost[9] i32, 172 reroll_cost i32 (shop.reroll_cost, grows x1.35 per reroll) */
#define BGYM_COLD_FIELDS \
  uint16_t deck[52]; uint8_t hand_play_count[12]; \
  uint8_t item_type[9]; uint8_t item_id[9]; uint8_t n_items; uint8_t deck_extra_n; \
  int32_t item_cost[9]; int32_t reroll_cost;

typedef struct BgymHot { BGYM_HOT_FIELDS } BgymHot;
typedef struct BgymCold { BGYM_COLD_FIELDS } BgymCold;
typedef struct BgymState { BGYM_HOT_FIELDS BGYM_COLD_FIELDS } BgymState;

/* ---- toggle record (32 B, device only) ------------------------------------------
 * Card-select toggles are ~75 % of the steps of a random-legal rollout, and a toggle reads and writes nothing but
 * the selection.  tog[n] is the dense array the select path runs on — the main pass of bgym_step reads 32 B and
 * writes 32 + 16 B per env instead of a 144-byte hot record and a 176-byte observation record:
 *   bytes  0..15  THE AUTHORITATIVE COPY of bytes 16..31 of the hot record (hand_n .. ep_len, "chunk 1").  A toggle
 *                 updates it here only: after a step, hot[i] bytes 16..31 may be stale for envs whose last actions were
 *                 toggles; every other byte of hot[i] is always current.  bgym_sync_state(BGYM_SYNC_TO_RECORDS) copies
 *                 the chunk back into the hot records (checkpoints, host copies, tests).
 *   bytes 16..31  a read-only summary of what else the select path needs: discards_left, cons_n (the PLAY-phase
 *                 action mask, balatro_env_2.py:1426-1445), guard = (ante > 100 || chips_scored > 10^9) (:619-623),
 *                 the Philox key word (fused random policy).  Rewritten by every pass that rewrites the hot record;
 *                 bgym_sync_state(BGYM_SYNC_FROM_RECORDS) rebuilds the whole toggle record from a hot record the
 *                 caller has written. */
typedef struct BgymTog {
  uint8_t hand_n; uint8_t hand_size; uint8_t sel_n; uint8_t highlight_mask; uint32_t sel_order;   /*  0 = hot bytes 16..23 */
  uint8_t face_down_mask; uint8_t phase; uint8_t round; uint8_t boss_type; uint32_t ep_len;       /*  8 = hot bytes 24..31 */
  uint8_t discards_left;  /* 16 */
  uint8_t cons_n;         /* 17 */
  uint8_t guard;          /* 18 ante > 100 || chips_scored > 1 000 000 000 */
  uint8_t _pad0;          /* 19 */
  uint32_t rng_seed;      /* 20 */
  uint8_t _pad1[8];       /* 24 */
} BgymTog;

/* ---- observation record (176 B) ---------------------------------------------
 * The 31 keys the reference actually emits (balatro_env_2.py:1488-1531), same dtypes
 * except selected_cards / face_down_cards (int8 here, platform int there) and the action
 * mask: the reference's int8[60] `action_mask` (:1522) travels as ONE 64-bit word,
 * action_mask_bits (bit a = action a legal) — 8 bytes instead of 60 on every record written
 * to HBM and copied over PCIe.  The Python layers expand it back into obs['action_mask']
 * (layout.mask_from_bits) wherever a caller asks for the reference's dict. */
typedef struct BgymObs {
  int8_t   hand[8];             /*   0 card code or -1                  */
  int8_t   selected_cards[8];   /*   8                                  */
  int8_t   face_down_cards[8];  /*  16                                  */
  int64_t  chips_scored;        /*  24                                  */
  int32_t  round_chips_scored;  /*  32                                  */
  float    progress_ratio;      /*  36 float32(min(2.0, round/max(1,needed))) */
  int32_t  mult;                /*  40 constant 1                       */
  int32_t  chips_needed;        /*  44                                  */
  int32_t  money;               /*  48                                  */
  int32_t  hands_played;        /*  52                                  */
  int32_t  best_hand_this_ante; /*  56                                  */
  int16_t  ante;                /*  60                                  */
  int16_t  shop_rerolls;        /*  62 state.shop_reroll_cost           */
  int16_t  joker_ids[10];       /*  64                                  */
  int16_t  consumables[5];      /*  84                                  */
  int16_t  shop_items[10];      /*  94                                  */
  int16_t  shop_costs[10];      /* 114                                  */
  int8_t   hand_levels[12];     /* 134                                  */
  int8_t   hand_size;           /* 146 len(hand_indexes)                */
  int8_t   deck_size;           /* 147                                  */
  int8_t   round;               /* 148                                  */
  int8_t   hands_left;          /* 149                                  */
  int8_t   discards_left;       /* 150                                  */
  int8_t   joker_count;         /* 151                                  */
  int8_t   joker_slots;         /* 152                                  */
  int8_t   consumable_count;    /* 153                                  */
  int8_t   consumable_slots;    /* 154                                  */
  int8_t   phase;               /* 155                                  */
  int8_t   boss_blind_active;   /* 156                                  */
  int8_t   boss_blind_type;     /* 157                                  */
  uint8_t  _pad0[2];            /* 158                                  */
  uint64_t action_mask_bits;    /* 160 bit a = action a legal (obs['action_mask'][a], :1522) */
  uint8_t  _pad1[8];            /* 168 (record stride 176 = 11 x 16 B)  */
} BgymObs;

/* ---- selection record (16 B, device only) ------------------------------------------
 * The two observation fields a card select changes, as a dense array sel[n]: THE AUTHORITATIVE COPY of
 * BgymObs.selected_cards and BgymObs.action_mask_bits.  Every step writes sel[i] for every env; the 176-byte record
 * obs[i] is rewritten only when another field changed (every action except a toggle or a rejected action; reset;
 * autoreset) — with all of its fields, these two included.  After a step, obs[i].selected_cards and
 * obs[i].action_mask_bits may therefore be stale for envs whose last actions were toggles; device-side consumers
 * (bgym_sample_actions, bgym_masked_sample) read the mask word from sel, and bgym_sync_obs(BGYM_SYNC_TO_RECORDS)
 * makes the records whole. */
typedef struct BgymSel {
  int8_t   selected_cards[8];   /* 0 = BgymObs.selected_cards   */
  uint64_t action_mask_bits;    /* 8 = BgymObs.action_mask_bits */
} BgymSel;

/* ---- per-step info record (32 B) --------------------------------------------
 * Fixed-width restatement of the numeric entries of the reference's info dict
 * (balatro_env_2.py:895-925). */
typedef struct BgymInfo {
  int64_t final_score;  /*  0 info['final_score'] (0 when no hand was scored)          */
  double  x_mult;       /*  8 score_breakdown['final_x_mult']                          */
  int32_t chips;        /* 16 score_breakdown['final_chips']                           */
  int32_t mult;         /* 20 score_breakdown['final_mult']                            */
  int8_t  hand_type;    /* 24 info['hand_type'] or -1                                  */
  uint8_t error_code;   /* 25 BGYM_ERR_*                                               */
  uint8_t flags;        /* 26 BGYM_F_*                                                 */
  uint8_t cards_played; /* 27 info['cards_played']                                     */
  int32_t base_score;   /* 28 UnifiedScorer result before steel/boss/retrigger         */
} BgymInfo;

/* ---- replay draws (256 B per env per step) -----------------------------------
 * Replay mode: the random draws the reference made during the same step, recorded
 * on the host (oracle/refenv.py::TapRandom), consumed by the kernel in order instead of Philox.
 *   u[]: results of random()/uniform(0,1) calls, in call order
 *   k[]: results of randint/choice/sample calls as 0-based indices into the
 *        population (sample contributes one entry per element drawn) */
typedef struct BgymDraws {
  double  u[24];
  uint8_t k[32];
  uint8_t n_u, n_k;    /* number of valid entries (for validation) */
  uint8_t _pad[30];
} BgymDraws;

/* ---- score_hands context (16 B per hand) -------------------------------------
 * The game_state entries the joker tables read (complete_joker_effects.py:39-50). */
typedef struct BgymScoreCtx {
  uint8_t hands_left;      /* Acrobat                                */
  uint8_t discards_left;   /* Mystic Summit, Banner                  */
  uint8_t deck_len;        /* Blue Joker: 2 * len(deck)              */
  uint8_t misprint[5];     /* replay: randint(0,23) result for the i-th Misprint joker */
  uint8_t bloodstone_bits; /* replay: bit c = (random() < 0.5) for scoring card c, first Bloodstone joker */
  uint8_t use_replay;      /* 0 = native Philox draws keyed by (seed, hand index), 1 = use the fields above */
  uint8_t _pad[6];
} BgymScoreCtx;

/* ---- device entry points ------------------------------------------------------ */
int bgym_abi_version(void);
const char* bgym_last_error(void);
/* process-wide tunables.  BGYM_OPT_SMALL_SLAB: slabs of at most `value` envs are stepped by the one-launch
 * kernel, larger ones by the main + gather passes (default 65536; 0 = always the multi-pass step).  Both give
 * identical results; the choice is a launch-latency / throughput trade. */
#define BGYM_OPT_SMALL_SLAB 1
int bgym_set_option(int option, int64_t value);
int bgym_device_count(void);

/* reset: for every env i with reset_mask == NULL || reset_mask[i] != 0 build a fresh episode
 * (balatro_env_2.py:505-558).  seeds[i] keys the native Philox stream.  decks52 != NULL
 * (n x 52 card codes) replays a supplied permutation (the reference's shuffle stream);
 * NULL = native Fisher-Yates from Philox.  Writes hot, tog, cold and — unless BGYM_FLAG_NO_OBS — obs and sel of the
 * envs it resets.  An env the mask leaves alone keeps its state (its toggle record is taken as current) and gets its
 * observation record and selection record re-emitted whole. */
int bgym_reset(BgymHot* hot, BgymTog* tog, BgymCold* cold, BgymObs* obs, BgymSel* sel, const uint8_t* reset_mask,
               const uint32_t* seeds, const uint8_t* decks52, int64_t n, int flags, void* stream);

/* step: one BalatroEnv.step per env (balatro_env_2.py:616-1064, 1174-1392).
 * draws == NULL -> native Philox mode.  info may be NULL. truncated is always 0.
 * `actions` is read (written instead with BGYM_FLAG_RANDOM_POLICY).
 * State: hot / tog / cold as described at BgymTog (a toggle touches tog only).  Observation: sel[i] is written for
 * every env, obs[i] for the envs whose other fields changed (see BgymSel); obs and sel must be the arrays the previous
 * reset / step of these envs wrote.  obs and sel may both be NULL with BGYM_FLAG_NO_OBS.
 * obs_dirty (n bytes, may be NULL): obs_dirty[i] is set to BGYM_OBS_DIRTY (| BGYM_OBS_DIRTY_SHOP when the record's shop block
 * — shop_items / shop_costs — can have changed) for every env whose observation RECORD this step rewrote; other bytes are
 * left alone.  The consumer of the deltas (bgym_pack_dirty_obs) clears them. */
enum { BGYM_OBS_DIRTY = 1, BGYM_OBS_DIRTY_SHOP = 2 };
int bgym_step(BgymHot* hot, BgymTog* tog, BgymCold* cold, int32_t* actions, const BgymDraws* draws, BgymObs* obs, BgymSel* sel,
              uint8_t* obs_dirty, double* reward, uint8_t* terminated, uint8_t* truncated, BgymInfo* info,
              int64_t n, int flags, void* stream);

/* Coherence between the record arrays and their device-only side arrays (BgymTog, BgymSel).
 *   BGYM_SYNC_TO_RECORDS    copy the authoritative fields into the records: hot[i] bytes 16..31 <- tog[i];
 *                           obs[i].selected_cards / .action_mask_bits <- sel[i].  After it the records are whole
 *                           (what checkpoints, host copies and the parity tests read).
 *   BGYM_SYNC_FROM_RECORDS  the caller has (re)written the records: rebuild tog[i] from hot[i] / sel[i] from obs[i]. */
enum { BGYM_SYNC_TO_RECORDS = 0, BGYM_SYNC_FROM_RECORDS = 1 };
int bgym_sync_state(BgymHot* hot, BgymTog* tog, int64_t n, int direction, void* stream);
int bgym_sync_obs(BgymObs* obs, BgymSel* sel, int64_t n, int direction, void* stream);

/* Observation deltas for a HOST mirror (the e2e path: ~65 B per env-step cross PCIe instead of 189 B).
 * A step rewrites obs[i] only for the envs it flags in obs_dirty; these two calls move exactly those records into a
 * mirror the GPU writes itself (zero-copy stores into pinned host memory: no host thread touches the records).
 * Mirror layout — chosen so that a record lands as ONE aligned 128-byte line, which is what makes zero-copy stores run at
 * the link's full rate (tools/exp/zc_probe.cu: 52 GB/s, against 35-42 GB/s for 176-byte records at a 176-byte stride):
 *   core[n]  BGYM_MIRROR_CORE_BYTES = 128 per env: the 16-byte chunks 0..5, 8, 9 of BgymObs, in that order
 *            (bytes 0..95 and 128..159: everything but the middle of the shop block and the mask word)
 *   shop[n]  BGYM_MIRROR_SHOP_BYTES = 32 per env: chunks 6, 7 (bytes 96..127: shop_items[1..9], shop_costs[0..6]) — all
 *            zero outside SHOP phase, so an env that was and stays in PLAY phase does not send them
 *   the mask word and selected_cards come from a copy of the selection array (BgymSel), reward / terminated from theirs.
 *   bgym_pack_dirty_obs     (after the step, on its stream) gathers the records flagged in obs_dirty, IN ASCENDING ENV ORDER
 *                           (zero-copy stores in address order run 20 % faster than in random order), into `staging`
 *                           (device): int32 count at byte 0, uint32 [cap] from byte 16 (env index, bit 31 = shop chunks
 *                           included), then from byte 16 + 4 * cap rounded up to 16 the first BGYM_OBS_DELTA_BYTES of each
 *                           record, [cap]; the flags it consumed are cleared.  obs_dirty == NULL stages every record (the
 *                           first fill of a mirror).  cap >= n is required; BGYM_DIRTY_STAGING_BYTES(cap) is the size of
 *                           `staging`, BGYM_DIRTY_SCRATCH_BYTES(n) that of `scratch` (device, block counts).
 *   bgym_scatter_dirty_obs  (any stream, once the pack has completed) writes staged record k to mirror_core[index] and,
 *                           when flagged, mirror_shop[index]; both may be device memory or PINNED HOST memory
 *                           (128- / 32-byte aligned). */
#define BGYM_MIRROR_CORE_BYTES 128
#define BGYM_MIRROR_SHOP_BYTES 32
#define BGYM_DIRTY_STAGING_BYTES(cap) (16 + (((size_t)(cap) * 4 + 15) & ~(size_t)15) + (size_t)(cap) * BGYM_OBS_DELTA_BYTES)
#define BGYM_DIRTY_SCRATCH_BYTES(n) ((((size_t)(n) + 4095) / 4096) * 4)
int bgym_pack_dirty_obs(const BgymObs* obs, uint8_t* obs_dirty, void* staging, void* scratch, int64_t cap, int64_t n, void* stream);
int bgym_scatter_dirty_obs(const void* staging, int64_t cap, void* mirror_core, void* mirror_shop, void* stream);

/* bgym_step keeps a few bytes per env of device scratch (work lists) per (device, stream) it is called on, sized for
 * the largest slab seen; this gives the scratch of `stream` on the current device back.  Call it once the stream's last
 * step has completed (bgym_vec_destroy does it for the handle's own stream). */
int bgym_release_stream(void* stream);

/* action mask as one 64-bit word per env (balatro_env_2.py:1426-1471) */
int bgym_action_mask(const BgymHot* hot, const BgymTog* tog, const BgymCold* cold, uint64_t* mask, int64_t n, void* stream);

/* uniform random legal action per env from its legal-action word (the policy the reference benchmarks are driven
 * with: random legal actions from obs['action_mask']).  mask_words points at env 0's word, mask_stride is the
 * distance in bytes between consecutive envs' words: (&sel->action_mask_bits, sizeof(BgymSel)) for the selection
 * array a step keeps current, (&obs->action_mask_bits, sizeof(BgymObs)) for whole observation records. */
int bgym_sample_actions(const uint64_t* mask_words, int64_t mask_stride, int32_t* actions, uint32_t seed, uint64_t step,
                        int64_t n, void* stream);

/* hand scoring: classify cards (balatro_game.py:40-93), per-card chip values
 * (balatro_env_2.py:287-325, cards.py:262-267), joker pipeline (unified_scoring.py:111-299).
 *   cards8  n x 8 card codes (0..51), entries >= n_cards[i] ignored
 *   mods8   n x 8 (enhancement | edition<<4 | seal<<8), NULL = no modifiers
 *   n_cards n, 1..8 (NULL = 5)
 *   jokers8 n x 8 joker ids, 0 = empty (NULL = no jokers)
 *   levels12 n x 12 hand levels (NULL = all 1)
 *   ctx     n (NULL = hands_left 4, discards_left 3, deck_len 52, native draws) */
int bgym_score_hands(const uint8_t* cards8, const uint16_t* mods8, const uint8_t* n_cards,
                     const uint8_t* jokers8, const uint8_t* levels12, const BgymScoreCtx* ctx,
                     uint8_t* hand_type, int32_t* chips, int32_t* mult, double* x_mult,
                     int64_t* score, int32_t* money, uint32_t seed, int64_t n, int flags, void* stream);

/* per-slab episode statistics: folds (reward, terminated) of one step into
 * stats[0]=episodes, [1]=sum return, [2]=sum length, [3]=steps, [4]=sum reward.
 * ret_acc / len_acc are n-sized per-env accumulators owned by the caller. */
int bgym_episode_stats(const double* reward, const uint8_t* terminated, double* ret_acc,
                       uint32_t* len_acc, double* stats, int64_t n, void* stream);

/* bgym_sample_actions with the step number kept on the device (*step_counter is read by the sampler and then
 * incremented): the form a CUDA graph can replay, since a replayed launch cannot take a new kernel argument */
int bgym_sample_actions_ctr(const uint64_t* mask_words, int64_t mask_stride, int32_t* actions, uint32_t seed,
                            uint64_t* step_counter, int64_t n, void* stream);

/* ---- on-device PPO rollout collection (config 5) ------------------ */
/* Observation records -> the dense input of the reference's BalatroFeaturesExtractor.forward
 * (train_balatro_agent.py:84-113): per env BGYM_FEATURE_DIM columns =
 *   [0,416)   one-hot of hand[8] over 52 card codes (empty slot = all zero)      :86-93
 *   [416,426) joker_ids[10] as floats                                            :98
 *   [426,447) chips_scored/1e6, chips_needed/1e5, progress_ratio, money/100, ante/10, round/3,
 *             hands_left/10, discards_left/5, hand_levels[12]/10, phase/3        :102-113
 *   [447]     0 (pad to a 16-byte multiple)
 * features: n x BGYM_FEATURE_DIM of dtype BGYM_DT_F32 or BGYM_DT_BF16, 16-byte aligned.
 * Reads none of the fields a BgymSel carries, so obs needs no bgym_sync_obs first (nor does bgym_policy_first_layer). */
int bgym_featurize(const BgymObs* obs, void* features, int64_t n, int dtype, void* stream);

/* First Linear + ReLU of the reference extractor's three sub-nets (train_balatro_agent.py:52-69: hand_net[0] 416 -> 256,
 * joker_net[0] 10 -> 128, game_state_net[0] 21 -> 64), computed straight from the observation records: the 8-hot hand
 * block makes hand_net[0] a sum of eight rows of its weight, so the 448-column feature matrix of bgym_featurize is never
 * materialised.  Weights are given TRANSPOSED ([in][out], bf16, 16-byte aligned): wt_hand [416][256], wt_joker [10][128],
 * wt_game [21][64]; bias = the three bias vectors back to back (448 floats).  out: n x 448 bf16 =
 * relu([hand 256 | joker 128 | game 64]), the inputs of hand_net[2] / joker_net[2] / game_state_net[2]. */
int bgym_policy_first_layer(const BgymObs* obs, const void* wt_hand, const void* wt_joker, const void* wt_game,
                            const float* bias, void* out, int64_t n, void* stream);

/* Masked categorical policy head: for each env, softmax over logits[60] restricted to the legal
 * actions of its mask word (mask_words / mask_stride as in bgym_sample_actions); draws one action by inverse CDF, returns its log-probability
 * and the entropy of the masked distribution (entropy may be NULL).  The uniform comes from
 * uniforms[i] when given, else from Philox keyed (seed, sample key) at counter (env_offset + i, step).
 * logits: n x 60, dtype BGYM_DT_F32 (16-byte aligned rows) or BGYM_DT_BF16 (8-byte aligned).
 * An env with no legal action gets action 0, logp 0, entropy 0. */
int bgym_masked_sample(const void* logits, int dtype, const uint64_t* mask_words, int64_t mask_stride, const float* uniforms,
                       uint32_t seed, uint64_t step, int64_t env_offset,
                       int32_t* actions, float* logp, float* entropy, int64_t n, void* stream);

/* GAE(gamma, lambda) over a [T, n] rollout stored time-major (what SB3's
 * RolloutBuffer.compute_returns_and_advantage computes for the reference's PPO,
 * train_balatro_agent.py:328-336): values is [T+1, n] (last row = bootstrap value),
 * dones[t] = 1 when step t ended its episode. */
int bgym_gae(const float* rewards, const float* values, const uint8_t* dones, float gamma, float lam,
             float* advantages, float* returns, int64_t T, int64_t n, void* stream);

/* ---- host-buffer entry points (what a non-torch caller binds) ------------------ */
typedef struct BgymVec BgymVec;
/* allocates device state/obs/outputs and pinned staging for n envs on `device` */
int bgym_vec_create(BgymVec** out, int64_t n, int device);
int bgym_vec_destroy(BgymVec* v);
/* host pointers; copies are done inside on the handle's stream and the call returns
 * after the results are in the host buffers */
int bgym_vec_reset_host(BgymVec* v, const uint32_t* seeds, const uint8_t* decks52, BgymObs* obs_out);
/* same, for the envs with reset_mask[i] != 0 only (the auto-reset of a host-driven vector env: SB3's
 * DummyVecEnv.step_wait resets finished envs one by one); untouched envs get their observation re-emitted */
int bgym_vec_reset_masked_host(BgymVec* v, const uint8_t* reset_mask, const uint32_t* seeds, const uint8_t* decks52,
                               BgymObs* obs_out);
/* flags: BGYM_FLAG_AUTORESET, BGYM_FLAG_GEN_*; BGYM_FLAG_RANDOM_POLICY is rejected (BGYM_E_ARG): `actions` is input only
 * here.  Every bgym_vec_* call runs on the handle's device and restores the caller's current device before returning. */
int bgym_vec_step_host(BgymVec* v, const int32_t* actions, const BgymDraws* draws, BgymObs* obs_out,
                       double* reward_out, uint8_t* terminated_out, uint8_t* truncated_out,
                       BgymInfo* info_out, int flags);
/* raw device pointers of the handle (for zero-copy consumers; any of the outputs may be NULL).  hot / obs are subject
 * to the staleness rules of BgymTog / BgymSel: read the toggle and selection arrays next to them. */
int bgym_vec_pointers(BgymVec* v, void** hot, void** tog, void** cold, void** obs, void** sel, void** reward, void** terminated);
/* copy state records to / from host (checkpointing: save_state/load_state,
 * balatro_env_2.py:1575-1615) */
int bgym_vec_get_state(BgymVec* v, BgymState* host_out);
int bgym_vec_set_state(BgymVec* v, const BgymState* host_in);

#ifdef __cplusplus
}
#endif
#endif /* BGYM_H */
