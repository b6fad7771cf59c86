// bgym_env.cuh — per-env game logic of the fused step kernel (device functions).
//
// One thread owns one env.  The env's hot record (144 B) and cold record (176 B: deck + shop) sit
// in shared memory, staged there by bulk async copies; the hot record is unpacked into registers
// (struct Hot), the cold record is accessed in shared memory because it is indexed dynamically.
// In every function below `rec` is the COLD record pointer; `hot` is the hot record pointer.
//
// Semantics follow the reference's *effective* behaviour; each function
// cites the reference lines it implements.
#pragma once
#include "bgym_device.cuh"

namespace bgym {

// offsets inside the cold record (BgymCold, include/bgym.h)
constexpr int OFF_DECK = 0, OFF_HPC = 104, OFF_ITEM_TYPE = 116, OFF_ITEM_ID = 125, OFF_N_ITEMS = 134,
              OFF_EXTRA_N = 135, OFF_ITEM_COST = 136, OFF_REROLL = 172;
// hot record bytes 136..143: deck_extra[4] (cards appended by Cryptid).  They are NOT part of struct Hot: only the
// consumable path reads or writes them, in the packed record; pack_hot leaves them alone.
constexpr int OFF_HOT_EXTRA = 136, MAX_DECK_EXTRA = 4;

enum { B_HOOK = 1, B_WALL, B_WHEEL, B_HOUSE, B_MARK, B_FISH, B_PSYCHIC, B_GOAD, B_WATER, B_WINDOW,
       B_MANACLE, B_EYE, B_MOUTH, B_PLANT, B_SERPENT, B_PILLAR, B_NEEDLE, B_HEAD, B_CLUB, B_TOOTH,
       B_FLINT, B_OXIDE, B_ARM, B_VIOLET, B_VERDANT, B_AMBER, B_CRIMSON, B_CERULEAN };

// ---------------------------------------------------------------------------------------------
// hot block in registers
// ---------------------------------------------------------------------------------------------
struct Hot {
  uint64_t hand, hand_code;          // 8 packed bytes each
  int hand_n, hand_size, sel_n, highlight;
  uint32_t sel_order;
  int face_down, phase, round, boss_type;
  int hands_left, discards_left, joker_n, cons_n;
  int joker_slots, cons_slots, n_magic, n_minimalist;
  int ante, jokers_sold, money, chips_needed;
  long long round_chips, chips_scored;
  int best_hand, hands_played_total, hands_played_ante;
  int boss_flags, boss_cards_required, boss_played_types, boss_hands_played, deck_n;
  uint64_t boss_played_cards, jokers, cons;
  uint32_t lv0, lv1, lv2;            // 12 hand levels, packed bytes
  int shop_reroll_state;
  uint32_t rng_seed, rng_ctr, ep_len, episode;
};

// CG: the record lies in global memory and is read once — straight from L2, no L1 line allocated
template <bool CG = false>
__device__ __forceinline__ void unpack_hot(const uint8_t* hot, Hot& h) {
  uint4 q0 = ldr128<CG>(hot), q1 = ldr128<CG>(hot + 16), q2 = ldr128<CG>(hot + 32), q3 = ldr128<CG>(hot + 48);
  uint4 q4 = ldr128<CG>(hot + 64), q5 = ldr128<CG>(hot + 80), q6 = ldr128<CG>(hot + 96), q7 = ldr128<CG>(hot + 112);
  uint4 q8 = ldr128<CG>(hot + 128);
  h.hand = u64_of(q0.x, q0.y); h.hand_code = u64_of(q0.z, q0.w);
  h.hand_n = q1.x & 0xFF; h.hand_size = (q1.x >> 8) & 0xFF; h.sel_n = (q1.x >> 16) & 0xFF; h.highlight = q1.x >> 24;
  h.sel_order = q1.y;
  h.face_down = q1.z & 0xFF; h.phase = (q1.z >> 8) & 0xFF; h.round = (q1.z >> 16) & 0xFF; h.boss_type = q1.z >> 24;
  h.ep_len = q1.w;
  h.joker_slots = q2.x & 0xFF; h.cons_slots = (q2.x >> 8) & 0xFF; h.n_magic = (q2.x >> 16) & 0xFF; h.n_minimalist = q2.x >> 24;
  h.ante = (int)(short)(q2.y & 0xFFFF); h.jokers_sold = (int)(short)(q2.y >> 16);
  h.money = (int)q2.z; h.chips_needed = (int)q2.w;
  h.round_chips = (long long)u64_of(q3.x, q3.y); h.chips_scored = (long long)u64_of(q3.z, q3.w);
  h.best_hand = (int)q4.x; h.hands_played_total = (int)q4.y;
  h.hands_played_ante = (int)(short)(q4.z & 0xFFFF); h.boss_flags = (q4.z >> 16) & 0xFF; h.boss_cards_required = q4.z >> 24;
  h.boss_played_types = q4.w & 0xFFFF; h.boss_hands_played = (q4.w >> 16) & 0xFF; h.deck_n = q4.w >> 24;
  h.boss_played_cards = u64_of(q5.x, q5.y); h.jokers = u64_of(q5.z, q5.w);
  h.cons = u64_of(q6.x, q6.y); h.lv0 = q6.z; h.lv1 = q6.w;
  h.lv2 = q7.x; h.shop_reroll_state = (int)q7.y; h.rng_seed = q7.z; h.rng_ctr = q7.w;
  h.hands_left = q8.x & 0xFF; h.discards_left = (q8.x >> 8) & 0xFF; h.joker_n = (q8.x >> 16) & 0xFF; h.cons_n = q8.x >> 24;
  h.episode = q8.y;
}

// bytes 16..31 of the hot record: hand_n | hand_size | sel_n | highlight, sel_order, face_down | phase | round | boss,
// ep_len — everything a SELECT toggle changes
__device__ __forceinline__ uint4 hot_chunk1(const Hot& h) {
  uint4 q;
  q.x = (h.hand_n & 0xFF) | ((h.hand_size & 0xFF) << 8) | ((h.sel_n & 0xFF) << 16) | ((uint32_t)(h.highlight & 0xFF) << 24);
  q.y = h.sel_order;
  q.z = (h.face_down & 0xFF) | ((h.phase & 0xFF) << 8) | ((h.round & 0xFF) << 16) | ((uint32_t)(h.boss_type & 0xFF) << 24);
  q.w = h.ep_len;
  return q;
}

template <bool CG = false>
__device__ __forceinline__ void pack_hot(uint8_t* hot, const Hot& h) {
  uint4 q;
  q.x = (uint32_t)h.hand; q.y = (uint32_t)(h.hand >> 32); q.z = (uint32_t)h.hand_code; q.w = (uint32_t)(h.hand_code >> 32);
  str128<CG>(hot, q);
  str128<CG>(hot + 16, hot_chunk1(h));
  q.x = (h.joker_slots & 0xFF) | ((h.cons_slots & 0xFF) << 8) | ((h.n_magic & 0xFF) << 16) | ((uint32_t)(h.n_minimalist & 0xFF) << 24);
  q.y = (h.ante & 0xFFFF) | ((uint32_t)(h.jokers_sold & 0xFFFF) << 16);
  q.z = (uint32_t)h.money; q.w = (uint32_t)h.chips_needed;
  str128<CG>(hot + 32, q);
  q.x = (uint32_t)h.round_chips; q.y = (uint32_t)((uint64_t)h.round_chips >> 32);
  q.z = (uint32_t)h.chips_scored; q.w = (uint32_t)((uint64_t)h.chips_scored >> 32);
  str128<CG>(hot + 48, q);
  q.x = (uint32_t)h.best_hand; q.y = (uint32_t)h.hands_played_total;
  q.z = (h.hands_played_ante & 0xFFFF) | ((h.boss_flags & 0xFF) << 16) | ((uint32_t)(h.boss_cards_required & 0xFF) << 24);
  q.w = (h.boss_played_types & 0xFFFF) | ((h.boss_hands_played & 0xFF) << 16) | ((uint32_t)(h.deck_n & 0xFF) << 24);
  str128<CG>(hot + 64, q);
  q.x = (uint32_t)h.boss_played_cards; q.y = (uint32_t)(h.boss_played_cards >> 32);
  q.z = (uint32_t)h.jokers; q.w = (uint32_t)(h.jokers >> 32);
  str128<CG>(hot + 80, q);
  q.x = (uint32_t)h.cons; q.y = (uint32_t)(h.cons >> 32); q.z = h.lv0; q.w = h.lv1;
  str128<CG>(hot + 96, q);
  q.x = h.lv2; q.y = (uint32_t)h.shop_reroll_state; q.z = h.rng_seed; q.w = h.rng_ctr;
  str128<CG>(hot + 112, q);
  q.x = (h.hands_left & 0xFF) | ((h.discards_left & 0xFF) << 8) | ((h.joker_n & 0xFF) << 16) | ((uint32_t)(h.cons_n & 0xFF) << 24);
  q.y = h.episode;
  str64<CG>(hot + 128, make_uint2(q.x, q.y));   // bytes 136..143 (deck_extra) stay as they are
}
__device__ __forceinline__ void hot_clear_extra(uint8_t* hot) { *reinterpret_cast<uint2*>(hot + OFF_HOT_EXTRA) = make_uint2(0u, 0u); }

// ---------------------------------------------------------------------------------------------
// toggle record (BgymTog, 32 B, include/bgym.h): bytes 0..15 = THE copy of hot bytes 16..31 that card toggles update,
// bytes 16..31 = the rest of what the select path reads (discards_left, cons_n, guard, Philox key word)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void set_chunk1(Hot& h, const uint4 q1) {
  h.hand_n = q1.x & 0xFF; h.hand_size = (q1.x >> 8) & 0xFF; h.sel_n = (q1.x >> 16) & 0xFF; h.highlight = q1.x >> 24;
  h.sel_order = q1.y;
  h.face_down = q1.z & 0xFF; h.phase = (q1.z >> 8) & 0xFF; h.round = (q1.z >> 16) & 0xFF; h.boss_type = q1.z >> 24;
  h.ep_len = q1.w;
}
__device__ __forceinline__ bool guard_pending(const Hot& h) { return h.ante > 100 || h.chips_scored > 1000000000LL; }   // :619-623
__device__ __forceinline__ uint4 tog_summary(const Hot& h) {
  uint4 q;
  q.x = (h.discards_left & 0xFF) | ((h.cons_n & 0xFF) << 8) | (guard_pending(h) ? 0x10000u : 0u);
  q.y = h.rng_seed; q.z = 0; q.w = 0;
  return q;
}
// the fields of Hot the select path reads (toggle, PLAY-phase mask, routing, guard, fused policy), from a toggle record
// alone; every other field of h stays unset.  The guard travels as one flag: it is put back as an ante that trips it.
__device__ __forceinline__ void unpack_tog(const uint4 t0, const uint4 t1, Hot& h) {
  set_chunk1(h, t0);
  h.discards_left = t1.x & 0xFF; h.cons_n = (t1.x >> 8) & 0xFF;
  h.ante = ((t1.x >> 16) & 0xFF) ? 101 : 1; h.chips_scored = 0;
  h.rng_seed = t1.y;
}
// a whole env for the list / reset tiles: the hot record with its toggle-owned chunk taken from the toggle record
template <bool CG = false>
__device__ __forceinline__ void load_hot(const uint8_t* hot, const uint8_t* tog, Hot& h) {
  unpack_hot<CG>(hot, h);
  set_chunk1(h, __ldcg(reinterpret_cast<const uint4*>(tog)));
}
template <bool CG = false>
__device__ __forceinline__ void store_tog(uint8_t* tog, const Hot& h) {
  str128<CG>(tog, hot_chunk1(h));
  str128<CG>(tog + 16, tog_summary(h));
}
template <bool CG = false>
__device__ __forceinline__ void store_hot(uint8_t* hot, uint8_t* tog, const Hot& h) { pack_hot<CG>(hot, h); store_tog<CG>(tog, h); }

__device__ __forceinline__ int hand_level(const Hot& h, int ht) {
  uint32_t w = ht < 4 ? h.lv0 : (ht < 8 ? h.lv1 : h.lv2);
  return (w >> (8 * (ht & 3))) & 0xFF;
}
__device__ __forceinline__ void bump_hand_level(Hot& h, int ht) {  // state.hand_levels[..] += 1, u8 saturating
  int lv = hand_level(h, ht);
  if (lv >= 255) return;
  uint32_t inc = 1u << (8 * (ht & 3));
  if (ht < 4) h.lv0 += inc; else if (ht < 8) h.lv1 += inc; else h.lv2 += inc;
}

__device__ __forceinline__ int deck16(const uint8_t* rec, int idx) {
  return *reinterpret_cast<const uint16_t*>(rec + OFF_DECK + 2 * idx);
}
__device__ __forceinline__ void set_deck16(uint8_t* rec, int idx, int v) {
  *reinterpret_cast<uint16_t*>(rec + OFF_DECK + 2 * idx) = (uint16_t)v;
}
__device__ __forceinline__ bool owns_joker(const Hot& h, int id) {
  // byte-wise equality over the 8 packed joker ids (id != 0)
  uint64_t x = h.jokers ^ (0x0101010101010101ull * (uint64_t)id);
  return (((x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull) != 0);
}

// ---------------------------------------------------------------------------------------------
// action mask, balatro_env_2.py:1426-1471
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t action_mask(const Hot& h, const uint8_t* rec) {
  uint64_t m = 0;
  if (h.phase == BGYM_PHASE_PLAY) {
    m = ((1ull << min(h.hand_n, 8)) - 1) << BGYM_A_SELECT_BASE;
    if (h.sel_n > 0) m |= 1ull << BGYM_A_PLAY_HAND;
    if (h.sel_n > 0 && h.discards_left > 0) m |= 1ull << BGYM_A_DISCARD;
    m |= ((1ull << h.cons_n) - 1) << BGYM_A_USE_CONS_BASE;
  } else if (h.phase == BGYM_PHASE_SHOP) {
    // the nine cost words with three loads in flight together, the compares unrolled (as a loop over n_items — one
    // dependent load per item, run twice per tile — this was a fifth of the shop list kernel's stall samples)
    const int n_items = rec[OFF_N_ITEMS];
    const uint2 c01 = *reinterpret_cast<const uint2*>(rec + OFF_ITEM_COST);
    const uint4 c25 = *reinterpret_cast<const uint4*>(rec + OFF_ITEM_COST + 8), c69 = *reinterpret_cast<const uint4*>(rec + OFF_ITEM_COST + 24);
    const int cost[9] = {(int)c01.x, (int)c01.y, (int)c25.x, (int)c25.y, (int)c25.z, (int)c25.w, (int)c69.x, (int)c69.y, (int)c69.z};
    #pragma unroll
    for (int i = 0; i < 9; i++)
      if (i < n_items && h.money >= cost[i]) m |= 1ull << (BGYM_A_SHOP_BUY_BASE + i);
    if (h.money >= h.shop_reroll_state) m |= 1ull << BGYM_A_SHOP_REROLL;
    m |= 1ull << BGYM_A_SHOP_END;
    m |= ((1ull << h.joker_n) - 1) << BGYM_A_SELL_JOKER_BASE;
  } else if (h.phase == BGYM_PHASE_BLIND_SELECT) {
    m = 0xFull << BGYM_A_SELECT_BLIND_BASE;  // 45,46,47 + SKIP_BLIND 48
  }
  return m;
}

// ---------------------------------------------------------------------------------------------
// hand list operations
// ---------------------------------------------------------------------------------------------
// The three hand-list helpers below run over the eight hand slots with STATIC slot indices (fully unrolled, predicated):
// as loops over hand_n with a run-time byte position they were ~25 instructions per slot (64-bit variable shifts to
// extract and insert a byte) with one dependent deck load per trip; unrolled, a slot is ~8 instructions and the eight
// deck loads are in flight together.
// BalatroGame._draw_cards balatro_game.py:95-109: top up with the lowest deck indices not in hand
__device__ __forceinline__ void draw_cards(Hot& h) {
  int want = h.hand_size - h.hand_n;
  if (want <= 0) return;
  uint64_t in_hand = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) if (i < h.hand_n) in_hand |= 1ull << ((h.hand >> (8 * i)) & 0xFF);
  uint64_t avail = ~in_hand & ((h.deck_n >= 64) ? ~0ull : ((1ull << h.deck_n) - 1));
  const int n0 = h.hand_n;
#pragma unroll
  for (int sl = 0; sl < 8; sl++) {
    if (sl >= n0 && sl < n0 + want && avail) {      // (avail runs dry for good: the filled slots stay a prefix)
      const uint64_t idx = (uint64_t)(__ffsll((long long)avail) - 1);
      avail &= avail - 1;
      h.hand = (h.hand & ~(0xFFull << (8 * sl))) | (idx << (8 * sl));
      h.hand_n++;
    }
  }
}
// remove the hand slots in `slots` (bit set), keeping order
__device__ __forceinline__ void remove_slots(Hot& h, int slots) {
  uint64_t out = 0;
  int n = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i < h.hand_n && !((slots >> i) & 1)) {
      out |= ((h.hand >> (8 * i)) & 0xFF) << (8 * n);
      n++;
    }
  }
  h.hand = out | (n >= 8 ? 0ull : (~0ull << (8 * n)));     // empty slots read 0xFF
  h.hand_n = n;
}
// hand_code cache: card code of every hand slot (what obs['hand'] shows), refreshed whenever the
// hand list changes so that passes that only have the hot record can emit the observation
__device__ __forceinline__ void refresh_hand_codes(Hot& h, const uint8_t* rec) {
  uint64_t codes = ~0ull;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int idx = (int)((h.hand >> (8 * i)) & 0xFF);
    if (i < h.hand_n && idx < h.deck_n)
      codes = (codes & ~(0xFFull << (8 * i))) | ((uint64_t)c16_code(deck16(rec, idx)) << (8 * i));
  }
  h.hand_code = codes;
}

// ---------------------------------------------------------------------------------------------
// shop, shop.py:96-205 and balatro_env_2.py:1383-1392
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double shop_cost_mult(const Hot& h) {  // shop.py:105-109
  double m = c_pow_1_15[min(max(h.ante - 1, 0), 100)];
  if (h.n_magic > 0) m *= 0.9;
  return m;
}

__device__ __forceinline__ void shop_put(uint8_t* rec, int i, int type, int id, int cost) {
  rec[OFF_ITEM_TYPE + i] = (uint8_t)type;
  rec[OFF_ITEM_ID + i] = (uint8_t)id;
  *reinterpret_cast<int*>(rec + OFF_ITEM_COST + 4 * i) = cost;
}

// shop.py:112-139.  The joker pool is "ids with base_cost > 0 not owned", which is ids
// 1..BGYM_NUM_SHOP_JOKERS minus the owned ones, so pool position p maps to an id by skipping
// owned ids in increasing order.
__device__ void shop_generate_inventory(Hot& h, uint8_t* rec, Draws& rng) {
  double mult = shop_cost_mult(h);
  int third = BGYM_PACK_TAROT + rng.below(3);
  shop_put(rec, 0, BGYM_ITEM_PACK, BGYM_PACK_STANDARD, (int)(c_pack_cost[BGYM_PACK_STANDARD] * mult));
  shop_put(rec, 1, BGYM_ITEM_PACK, BGYM_PACK_JOKER, (int)(c_pack_cost[BGYM_PACK_JOKER] * mult));
  shop_put(rec, 2, BGYM_ITEM_PACK, third, (int)(c_pack_cost[third] * mult));
  // owned shop-eligible ids, sorted ascending (<= 8 entries)
  int owned[8], n_owned = 0;
  #pragma unroll 1
  for (int i = 0; i < h.joker_n; i++) {
    int id = byte_at(h.jokers, i);
    if (id >= 1 && id <= BGYM_NUM_SHOP_JOKERS) {
      int k = n_owned++;
      #pragma unroll 1
      while (k > 0 && owned[k - 1] > id) { owned[k] = owned[k - 1]; k--; }
      owned[k] = id;
    }
  }
  // distinct owned ids only (Ankh-style duplicates cannot occur in-env, but stay safe)
  int m = 0;
  #pragma unroll 1
  for (int i = 0; i < n_owned; i++) if (i == 0 || owned[i] != owned[i - 1]) owned[m++] = owned[i];
  n_owned = m;
  int pool = BGYM_NUM_SHOP_JOKERS - n_owned;
  int k = min(3, pool);
  int n = 3;
  int chosen[3];
  #pragma unroll 1
  for (int t = 0; t < k; t++) {
    int p;
    if (rng.tape) {
      p = rng.below(pool);  // replay: population index recorded from the reference
    } else {
      // native: t-th element of a uniform ordered sample without replacement
      p = rng.below(pool - t);
      // map to the p-th not-yet-chosen position (chosen kept sorted)
      #pragma unroll 1
      for (int a = 0; a < t; a++) if (chosen[a] <= p) p++;
    }
    // keep `chosen` sorted ascending for the skip logic above
    int c = t;
    #pragma unroll 1
    while (c > 0 && chosen[c - 1] > p) { chosen[c] = chosen[c - 1]; c--; }
    chosen[c] = p;
    int id = p + 1;
    #pragma unroll 1
    for (int a = 0; a < n_owned; a++) if (owned[a] <= id) id++;
    shop_put(rec, n++, BGYM_ITEM_JOKER, id, (int)(c_joker_cost[id] * mult));
  }
  int v = rng.below(2);
  shop_put(rec, n++, BGYM_ITEM_VOUCHER, v, (int)(c_voucher_cost[v] * mult));
  #pragma unroll 1
  for (int i = 0; i < 2; i++) {
    int c = rng.below(52);
    shop_put(rec, n++, BGYM_ITEM_CARD, c, BGYM_CARD_COST);
  }
  rec[OFF_N_ITEMS] = (uint8_t)n;
  #pragma unroll 1
  for (int i = n; i < 9; i++) shop_put(rec, i, 0, 0, 0);
}

__device__ __forceinline__ void generate_shop(Hot& h, uint8_t* rec, Draws& rng) {  // balatro_env_2.py:1383-1392
  *reinterpret_cast<int*>(rec + OFF_REROLL) = BGYM_REROLL_BASE;
  shop_generate_inventory(h, rec, rng);
  h.shop_reroll_state = (int)(BGYM_REROLL_BASE * shop_cost_mult(h));
}

// ---------------------------------------------------------------------------------------------
// round advance, balatro_env_2.py:1326-1381
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void boss_deactivate(Hot& h) {
  h.boss_type = 0; h.boss_flags = 0; h.boss_cards_required = 0; h.boss_played_types = 0;
  h.boss_hands_played = 0; h.boss_played_cards = 0;
}

__device__ void advance_round(Hot& h, uint8_t* rec, Draws& rng) {
  int gold = 0;
  #pragma unroll 1
  for (int i = 0; i < h.hand_n; i++) {
    int idx = byte_at(h.hand, i);
    if (idx < 52 && c16_enh(deck16(rec, idx)) == BGYM_ENH_GOLD) gold += 3;
  }
  h.money += gold;
  if (h.boss_type) {
    h.money += 5;  // BossBlind.money_reward, boss_blinds.py:60
    boss_deactivate(h);
    h.face_down = 0;
  }
  h.round_chips = 0; h.best_hand = 0; h.hands_played_ante = 0;
  if (h.round == 3) {
    h.ante += 1; h.round = 1;
    if (h.ante > 100) return;
  } else {
    h.round += 1;
  }
  h.money += 25 * h.round + (h.round == 3 ? 10 : 0);
  h.hands_left = 4; h.discards_left = 3;
  h.phase = BGYM_PHASE_SHOP;
  generate_shop(h, rec, rng);
}

// ---------------------------------------------------------------------------------------------
// consumables, balatro_env_2.py:1066-1172 + consumables.py:115-655
// Table-driven: each consumable id maps to an (op, arg) pair; the op set is small.
// ---------------------------------------------------------------------------------------------
enum ConsOp : uint8_t {
  CO_NONE = 0,
  CO_ENH_N,        // set enhancement `arg & 15` on the first `arg >> 4` targets (>= 1 target needed)
  CO_AFFECT_N,     // targets only "affected" (effect lost in the reference): first `arg` targets
  CO_STRENGTH,     // first 2 targets, affected only when rank < ace
  CO_DEATH,        // needs 2 targets, 2 affected
  CO_HERMIT, CO_TEMPERANCE,
  CO_WHEEL,        // 25% roll then edition
  CO_FOOL, CO_PRIESTESS, CO_EMPEROR, CO_JUDGEMENT,
  CO_RAISE_IF_TARGET,   // reference raises when a target is selected, fails otherwise
  CO_RAISE_IF_HAND,     // Sigil / Ouija: draws then raises (arg = population)
  CO_PLANET,       // arg = hand type
  CO_SEAL,         // arg = raw seal value stored (consumables.Seal numbering)
  CO_AURA, CO_WRAITH, CO_ECTOPLASM, CO_ANKH, CO_HEX, CO_SOUL, CO_BLACK_HOLE,
  CO_IMMOLATE, CO_CRYPTID
};
struct ConsRow { uint8_t op, arg; };

constexpr ConsRow cons_row_of(int cid) {
  int t = (cid >= 101 && cid <= 122) ? cid - 100 : cid;  // enum-style tarot names resolve the same way
  switch (t) {
    case 1: return {CO_FOOL, 0};
    case 2: return {CO_ENH_N, (2 << 4) | BGYM_ENH_LUCKY};
    case 3: return {CO_PRIESTESS, 0};
    case 4: return {CO_ENH_N, (2 << 4) | BGYM_ENH_MULT};
    case 5: return {CO_EMPEROR, 0};
    case 6: return {CO_ENH_N, (2 << 4) | BGYM_ENH_BONUS};
    case 7: return {CO_ENH_N, (1 << 4) | BGYM_ENH_WILD};
    case 8: return {CO_ENH_N, (1 << 4) | BGYM_ENH_STEEL};
    case 9: return {CO_STRENGTH, 0};
    case 10: return {CO_HERMIT, 0};
    case 11: return {CO_WHEEL, 0};
    case 12: return {CO_ENH_N, (1 << 4) | BGYM_ENH_GLASS};
    case 13: return {CO_RAISE_IF_TARGET, 0};
    case 14: return {CO_DEATH, 0};
    case 15: return {CO_TEMPERANCE, 0};
    case 16: return {CO_ENH_N, (1 << 4) | BGYM_ENH_GOLD};
    case 17: return {CO_ENH_N, (1 << 4) | BGYM_ENH_STONE};
    case 18: case 19: case 20: case 22: return {CO_AFFECT_N, 3};
    case 21: return {CO_JUDGEMENT, 0};
    // planets 30..41 -> hand type (balatro_env_2.py:1103-1116)
    case 30: return {CO_PLANET, 1}; case 31: return {CO_PLANET, 2}; case 32: return {CO_PLANET, 3};
    case 33: return {CO_PLANET, 4}; case 34: return {CO_PLANET, 5}; case 35: return {CO_PLANET, 6};
    case 36: return {CO_PLANET, 7}; case 37: return {CO_PLANET, 8}; case 38: return {CO_PLANET, 0};
    case 39: return {CO_PLANET, 9}; case 40: return {CO_PLANET, 10}; case 41: return {CO_PLANET, 11};
    // spectrals 50..67 (consumables.py:344-362)
    case 50: case 51: case 52: return {CO_RAISE_IF_TARGET, 0};
    case 53: return {CO_SEAL, 3};   // Talisman: consumables.Seal.GOLD == 3
    case 54: return {CO_AURA, 0};
    case 55: return {CO_WRAITH, 0};
    case 56: return {CO_RAISE_IF_HAND, 4};
    case 57: return {CO_RAISE_IF_HAND, 13};
    case 58: return {CO_ECTOPLASM, 0};
    case 60: return {CO_ANKH, 0};
    case 61: return {CO_SEAL, 1};   // Deja Vu: consumables.Seal.RED == 1
    case 62: return {CO_HEX, 0};
    case 63: return {CO_SEAL, 2};   // Trance: consumables.Seal.BLUE == 2
    case 64: return {CO_SEAL, 4};   // Medium: PURPLE == 4
    case 66: return {CO_SOUL, 0};
    case 67: return {CO_BLACK_HOLE, 0};
    case 59: return {CO_IMMOLATE, 0};
    case 65: return {CO_CRYPTID, 0};
  }
  return {CO_NONE, 0};
}
// the same as a 128-entry table in global memory (built at compile time): the lanes of a consumable tile hold ~20 different
// ids, and the switch above ran once per distinct id (ncu: 3 % of the list kernel's instructions at 3.8 active lanes)
struct ConsTable { uint16_t v[128]; };
constexpr ConsTable make_cons_table() {
  ConsTable t{};
  for (int i = 0; i < 128; i++) { const ConsRow r = cons_row_of(i); t.v[i] = (uint16_t)(r.op | (r.arg << 8)); }
  return t;
}
__device__ const ConsTable g_cons_table = make_cons_table();
__device__ __forceinline__ ConsRow cons_row(int cid) {
  const uint32_t v = (cid >= 0 && cid < 128) ? g_cons_table.v[cid] : 0u;
  return {(uint8_t)(v & 0xFF), (uint8_t)(v >> 8)};
}

__device__ __forceinline__ void cons_append(Hot& h, int cid) {
  if (h.cons_n < 8) { h.cons = with_byte(h.cons, h.cons_n, cid); h.cons_n++; }
}
__device__ __forceinline__ void cons_pop(Hot& h, int idx) {
  uint64_t lo = h.cons & ((1ull << (8 * idx)) - 1);
  uint64_t hi = (idx >= 7) ? 0 : (h.cons >> (8 * (idx + 1))) << (8 * idx);
  h.cons = lo | hi;
  h.cons_n--;
}

// Immolate (consumables.py:520-532): random.sample(deck, min(5, len)) then deck.remove(card) for each — the deck list
// really shrinks, hand_indexes are not re-mapped, card_states stay keyed by deck INDEX.  The deck is always
// [surviving original cards, in order] ++ [cards appended by Cryptid, in order]; list.remove takes the first EQUAL
// element: an original card (cards.Card, rank+suit equality, :112-115) is only equal to itself, an appended card
// (consumables.Card dataclass) is equal to any appended card with the same rank, suit and copy-time modifiers.
// Two halves.  immolate_sample (the lane that uses the card, inside use_consumable): the draws and which list entries
// go — a bit set over deck positions.  immolate_compact (the WHOLE WARP, called by the tile at a converged point for every
// lane that has such a set): the list compaction, lane = deck position.  One lane walking the 52 positions alone was
// 1 900 dependent warp-instructions that the other 31 lanes of the tile waited for (ncu: 19 % of the consumable list
// kernel's instructions at 1.1 active lanes); the cooperative form is ~80.
__device__ __noinline__ int immolate_sample(const Hot& h, const uint8_t* hotrec, const uint8_t* rec, Draws& rng, uint64_t* removed_out) {
  const int n = h.deck_n, n_ex = rec[OFF_EXTRA_N], n_orig = n - n_ex;
  const uint16_t* ex = reinterpret_cast<const uint16_t*>(hotrec + OFF_HOT_EXTRA);
  const int k = min(5, n);
  uint64_t sampled = 0, removed = 0;
#pragma unroll 1
  for (int t = 0; t < k; t++) {
    int idx;
    if (rng.tape) idx = rng.below(n);            // replay: population index recorded from the reference
    else {                                       // native: t-th element of a uniform sample without replacement
      int p = rng.below(n - t);
      idx = select_bit64(~sampled & ((n >= 64) ? ~0ull : ((1ull << n) - 1)), p);
    }
    sampled |= 1ull << idx;
    int victim = idx;
    if (idx >= n_orig) {
      const int id = ex[idx - n_orig];
#pragma unroll 1
      for (int j = n_ex - 1; j >= 0; j--) if (!((removed >> (n_orig + j)) & 1) && ex[j] == id) victim = n_orig + j;
    }
    removed |= 1ull << victim;
  }
  *removed_out = removed;
  return k;
}
// compact: codes move down with the cards, modifiers stay with the deck index.  Whole warp; `removed` != 0 marks the
// lanes whose env used Immolate this step (h, hotrec, rec are that lane's).  Lane l serves list positions l and l + 32
// of one env at a time: all reads first, then every surviving card's code goes to its new position.
__device__ __forceinline__ void immolate_compact(uint64_t removed, Hot& h, uint8_t* hotrec, uint8_t* rec, int lane) {
  uint32_t need = __ballot_sync(0xffffffffu, removed != 0);
  if (!need) return;
  __syncwarp();     // the lane's own accesses to its records are ordered before the warp reads and rewrites them
  const unsigned long long my_rec = reinterpret_cast<unsigned long long>(rec), my_hot = reinterpret_cast<unsigned long long>(hotrec);
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    const uint64_t rm = __shfl_sync(0xffffffffu, removed, src);
    uint8_t* r = reinterpret_cast<uint8_t*>(__shfl_sync(0xffffffffu, my_rec, src));
    uint16_t* ex = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(__shfl_sync(0xffffffffu, my_hot, src)) + OFF_HOT_EXTRA);
    const int n = __shfl_sync(0xffffffffu, h.deck_n, src);
    const uint64_t bpc = __shfl_sync(0xffffffffu, h.boss_played_cards, src);
    const int n_ex = r[OFF_EXTRA_N], n_orig = n - n_ex;
    const uint64_t keep = ~rm & ((1ull << n) - 1);                       // n <= 56
    const uint64_t keep_ex = keep & ~((1ull << n_orig) - 1);
    const int w_total = __popcll(keep), w_ex_total = __popcll(keep_ex);
    int code[2], id[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int i = lane + 32 * q;
      id[q] = (i >= n_orig && i < n) ? ex[i - n_orig] : 0;
      code[q] = i < n_orig ? c16_code(deck16(r, i)) : (id[q] & 63);
    }
    __syncwarp();
    uint64_t pillar = 0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int i = lane + 32 * q;
      if ((keep >> i) & 1) {
        const uint64_t below = (1ull << i) - 1;
        const int w = __popcll(keep & below);
        if (w < 52) set_deck16(r, w, (deck16(r, w) & ~63) | code[q]);
        if (i >= n_orig) ex[__popcll(keep_ex & below)] = (uint16_t)id[q];
        pillar |= ((bpc >> i) & 1ull) << w;
      }
      if (i >= w_total && i < 52) set_deck16(r, i, deck16(r, i) & ~63);
    }
    if (lane >= w_ex_total && lane < MAX_DECK_EXTRA) ex[lane] = 0;
    if (lane == 0) r[OFF_EXTRA_N] = (uint8_t)w_ex_total;
    const uint32_t plo = __reduce_or_sync(0xffffffffu, (uint32_t)pillar), phi = __reduce_or_sync(0xffffffffu, (uint32_t)(pillar >> 32));
    if (lane == src) { h.deck_n = w_total; h.boss_played_cards = u64_of(plo, phi); }
    __syncwarp();
  }
}

// *immolate_removed: set by Immolate to the deck positions it destroys (0 otherwise); the caller's tile runs
// immolate_compact() on it once the warp has converged
__device__ double use_consumable(Hot& h, uint8_t* hotrec, uint8_t* rec, int cidx, Draws& rng, int& err, int& terminated,
                                 uint64_t* immolate_removed) {
  const unsigned entered = __activemask();   // the lanes that walk this call together
  *immolate_removed = 0;
  int cid = byte_at(h.cons, cidx);
  ConsRow row = cons_row(cid);
  // targets: selected cards in selection order (balatro_env_2.py:1074-1083)
  int tgt[8], nT = 0;
  #pragma unroll 1
  for (int k = 0; k < h.sel_n; k++) {
    int sl = nib_at(h.sel_order, k);
    if (sl < h.hand_n) { int idx = byte_at(h.hand, sl); if (idx < h.deck_n) tgt[nT++] = idx; }
  }
  bool success = false, raise = false, unsupported = false;
  int money_gained = 0, planet_ht = -1, n_affected = 0, n_jokers_created = 0, add_joker = 0;
  int items[2], n_items = 0, hand_size_change = 0, n_created = 0, n_destroyed = 0;
  switch (row.op) {
    case CO_ENH_N: {
      int cnt = min(nT, row.arg >> 4);
      #pragma unroll 1
      for (int i = 0; i < cnt; i++) { int c = deck16(rec, tgt[i]); set_deck16(rec, tgt[i], (c & ~(15 << 6)) | ((row.arg & 15) << 6)); }
      n_affected = cnt; success = nT > 0;
      break;
    }
    case CO_AFFECT_N: n_affected = min(nT, (int)row.arg); success = nT > 0; break;
    case CO_STRENGTH:
      #pragma unroll 1
      for (int i = 0; i < min(nT, 2); i++) n_affected += (c16_code(deck16(rec, tgt[i])) >> 2) < 12;
      success = nT > 0;
      break;
    case CO_DEATH: if (nT >= 2) { n_affected = 2; success = true; } break;
    case CO_HERMIT: money_gained = min(h.money, 20); success = true; break;
    case CO_TEMPERANCE: money_gained = min(5 * h.joker_n, 50); success = true; break;
    case CO_WHEEL:
      if (nT > 0 && rng.u01() < 0.25) {
        int ed = BGYM_ED_FOIL + rng.below(3);
        int c = deck16(rec, tgt[0]); set_deck16(rec, tgt[0], (c & ~(7 << 10)) | (ed << 10));
        n_affected = 1; success = true;
      }
      break;
    case CO_FOOL: {  // appends to the aliased list without a slot check
      int copied = byte_at(h.cons, rng.below(h.cons_n));
      cons_append(h, copied); items[n_items++] = copied; success = true;
      break;
    }
    case CO_PRIESTESS:
      #pragma unroll 1
      for (int i = 0; i < 2; i++) {
        int p = BGYM_CONS_PLANET_BASE + rng.below(9);
        if (h.cons_n < h.cons_slots) { cons_append(h, p); items[n_items++] = p; }
      }
      success = true;
      break;
    case CO_EMPEROR:
      #pragma unroll 1
      for (int i = 0; i < 2; i++)
        if (h.cons_n < h.cons_slots) {
          int t = BGYM_CONS_ENUMSTYLE_BASE + 1 + rng.below(22);
          cons_append(h, t); items[n_items++] = t;
        }
      success = true;
      break;
    case CO_JUDGEMENT: {
      int p = BGYM_CONS_PLANET_BASE + rng.below(9);
      if (h.cons_n < h.cons_slots) { cons_append(h, p); items[n_items++] = p; }
      success = true;
      break;
    }
    case CO_RAISE_IF_TARGET: raise = nT >= 1; break;
    case CO_RAISE_IF_HAND: if (h.hand_n > 0) { (void)rng.below(row.arg); raise = true; } break;
    case CO_PLANET: planet_ht = row.arg; success = true; break;
    case CO_SEAL:
      if (nT >= 1) { int c = deck16(rec, tgt[0]); set_deck16(rec, tgt[0], (c & ~(7 << 13)) | (row.arg << 13)); n_affected = 1; success = true; }
      break;
    case CO_AURA:
      if (nT >= 1) {
        int ed = BGYM_ED_FOIL + rng.below(3);
        int c = deck16(rec, tgt[0]); set_deck16(rec, tgt[0], (c & ~(7 << 10)) | (ed << 10));
        n_affected = 1; success = true;
      }
      break;
    case CO_WRAITH:
      if (h.joker_n < h.joker_slots) {
        // consumables.py:479-481; 'Drivers License' (index 4) is not a library name -> nothing added
        int j = rng.below(14);
        add_joker = j < 4 ? BGYM_J_INVISIBLE_JOKER + j : (j == 4 ? 0 : BGYM_J_CARTOMANCER + (j - 5));
        n_jokers_created = 1; hand_size_change = -1; success = true;
      }
      break;
    case CO_ECTOPLASM: if (h.joker_n > 0) { hand_size_change = -1; success = true; } break;
    case CO_ANKH: if (h.joker_n > 0) { (void)rng.below(h.joker_n); n_jokers_created = 1; success = true; } break;
    case CO_HEX: if (h.joker_n > 0) { (void)rng.below(h.joker_n); success = true; } break;
    case CO_SOUL:
      if (h.joker_n < h.joker_slots) { add_joker = BGYM_J_CANIO + rng.below(5); n_jokers_created = 1; success = true; }
      break;
    case CO_BLACK_HOLE: success = true; break;
    case CO_IMMOLATE: n_destroyed = immolate_sample(h, hotrec, rec, rng, immolate_removed); money_gained = 20; success = true; break;
    case CO_CRYPTID:   // consumables.py:582-592: two copies of the first target are appended to the deck list
      if (nT >= 1) {
        const int n_ex = rec[OFF_EXTRA_N];
        if (n_ex + 2 > MAX_DECK_EXTRA || h.deck_n + 2 > 56) { unsupported = true; break; }   // capacity, include/bgym.h
        uint16_t* ex = reinterpret_cast<uint16_t*>(hotrec + OFF_HOT_EXTRA);
        const int id = deck16(rec, tgt[0]);
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
          ex[n_ex + q] = (uint16_t)id;
          if (h.deck_n < 52) set_deck16(rec, h.deck_n, (deck16(rec, h.deck_n) & ~63) | (id & 63));
          h.deck_n++;
        }
        rec[OFF_EXTRA_N] = (uint8_t)(n_ex + 2);
        n_created = 2; success = true;
      }
      break;
    default: break;
  }
  // every lane took its own case: meet again here, so that the bookkeeping below runs once for the tile and not once
  // per case (ncu: lines :609-:625 at 2 active lanes, 10 % of the list kernel's instructions)
  __syncwarp(entered);
  if (raise) {  // SafeBalatroEnv convention, train_balatro_fixed.py:262-269
    err = BGYM_ERR_REF_EXCEPTION; terminated = 1;
    return -100.0;
  }
  double reward = 0.0;
  if (unsupported) { err = BGYM_ERR_UNSUPPORTED; reward = -1.0; }
  else if (success) {
    cons_pop(h, cidx);
    if (money_gained > 0) { h.money += money_gained; reward += money_gained / 10.0; }
    if (planet_ht >= 0) { bump_hand_level(h, planet_ht); reward += 10.0; }
    if (n_affected > 0) reward += n_affected * 2.0;
    if (n_created > 0) reward += n_created * 3.0;          // :1140-1141
    if (n_destroyed > 0) reward += n_destroyed * 1.0;      // :1143-1144
    if (n_jokers_created > 0) {
      if (add_joker != 0 && h.joker_n < h.joker_slots && h.joker_n < 8) { h.jokers = with_byte(h.jokers, h.joker_n, add_joker); h.joker_n++; }
      reward += n_jokers_created * 15.0;
    }
    if (n_items > 0) {
      #pragma unroll 1
      for (int i = 0; i < n_items; i++) if (h.cons_n < h.cons_slots) cons_append(h, items[i]);
      reward += n_items * 5.0;
    }
    if (hand_size_change) h.hand_size = max(h.hand_size + hand_size_change, 0);
  } else {
    reward = -1.0; err = BGYM_ERR_CONSUMABLE_FAILED;
  }
  h.sel_n = 0; h.sel_order = 0;
  return reward;
}

// ---------------------------------------------------------------------------------------------
// reset, balatro_env_2.py:505-558 (UnifiedGameState defaults :166-211)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t next_episode_seed(uint32_t seed) {
  uint32_t x = seed + 0x9E3779B9u;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x ? x : 1u;
}

// ---------------------------------------------------------------------------------------------
// synthetic-state generator of BASELINE configs[2]/[3] (BGYM_FLAG_GEN_C3 / BGYM_FLAG_GEN_CONS; the law is
// specified in include/bgym.h, the reference-side counterpart is the injection recipe of SURVEY Appendix E)
// ---------------------------------------------------------------------------------------------
// modifier bits (enhancement << 6 | edition << 10 | seal << 13) of card k = suit * 13 + rank - 2
__device__ __forceinline__ int gen_mods_from_words(uint32_t w0, uint32_t w1) {
  const int enh = (w0 >> 30) == 0u ? 1 + (int)((w0 >> 27) & 7u) : 0;
  const int e = (int)(((w0 & 0x07FFFFFFu) * 30ull) >> 27), sl = (int)__umulhi(w1, 40u);
  return (enh << 6) | ((e < 3 ? 1 + e : 0) << 10) | ((sl < 4 ? 1 + sl : 0) << 13);
}
__device__ __forceinline__ int gen_card_mods(uint32_t seed, int k) {
  const uint4 b = philox4x32_10((uint32_t)(k >> 1), 0, 0, 0, seed, BGYM_GEN_KEY1);   // two cards per block
  return (k & 1) ? gen_mods_from_words(b.z, b.w) : gen_mods_from_words(b.x, b.y);
}
__device__ __forceinline__ int gen_cons_id(uint32_t w) {
  const int i = (int)__umulhi(w, 52u);
  return i < 22 ? BGYM_CONS_TAROT_BASE + i : (i < 34 ? BGYM_CONS_PLANET_BASE + (i - 22) : BGYM_CONS_SPECTRAL_BASE + (i - 34));
}
// jokers (five distinct shop-eligible ids, draw order) and, with BGYM_FLAG_GEN_CONS, both consumable slots
__device__ __noinline__ void gen_hot_fields(uint32_t seed, int flags, uint64_t* jokers_out, uint32_t* cons_out) {
  const uint4 a = philox4x32_10(64u, 0, 0, 0, seed, BGYM_GEN_KEY1);
  const uint4 b = philox4x32_10(65u, 0, 0, 0, seed, BGYM_GEN_KEY1);
  const uint32_t w[5] = {a.x, a.y, a.z, a.w, b.x};
  int chosen[5];   // pool positions picked so far, ascending
  uint64_t jk = 0;
#pragma unroll
  for (int t = 0; t < 5; t++) {
    int p = (int)__umulhi(w[t], (uint32_t)(BGYM_NUM_SHOP_JOKERS - t));
#pragma unroll
    for (int q = 0; q < t; q++) if (chosen[q] <= p) p++;      // p-th position not picked yet
    jk |= (uint64_t)(p + 1) << (8 * t);
    chosen[t] = p;                                            // one bubble pass keeps `chosen` ascending
#pragma unroll
    for (int q = t; q > 0; q--) {
      const int lo = min(chosen[q - 1], chosen[q]), hi = max(chosen[q - 1], chosen[q]);
      chosen[q - 1] = lo; chosen[q] = hi;
    }
  }
  *jokers_out = jk;
  *cons_out = (flags & BGYM_FLAG_GEN_CONS) ? ((uint32_t)gen_cons_id(b.y) | ((uint32_t)gen_cons_id(b.z) << 8)) : 0u;
}
__device__ __forceinline__ void gen_hot(Hot& h, uint32_t seed, int flags) {
  uint64_t jk; uint32_t cs;
  gen_hot_fields(seed, flags, &jk, &cs);
  h.jokers = jk; h.joker_n = 5;
  if (flags & BGYM_FLAG_GEN_CONS) { h.cons = cs; h.cons_n = 2; }
}

// hot block of a fresh episode (UnifiedGameState defaults)
__device__ __forceinline__ void reset_hot(Hot& h, uint32_t seed) {
  h.hand = ~0ull; h.hand_code = ~0ull;
  h.hand_n = 0; h.hand_size = 8; h.sel_n = 0; h.highlight = 0; h.sel_order = 0;
  h.face_down = 0; h.phase = BGYM_PHASE_BLIND_SELECT; h.round = 1; h.boss_type = 0;
  h.hands_left = 4; h.discards_left = 3; h.joker_n = 0; h.cons_n = 0;
  h.joker_slots = 5; h.cons_slots = 2; h.n_magic = 0; h.n_minimalist = 0;
  h.ante = 1; h.jokers_sold = 0; h.money = 4; h.chips_needed = 300;
  h.round_chips = 0; h.chips_scored = 0; h.best_hand = 0; h.hands_played_total = 0; h.hands_played_ante = 0;
  h.boss_flags = 0; h.boss_cards_required = 0; h.boss_played_types = 0; h.boss_hands_played = 0; h.deck_n = 52;
  h.boss_played_cards = 0; h.jokers = 0; h.cons = 0;
  h.lv0 = h.lv1 = h.lv2 = 0x01010101u;
  h.shop_reroll_state = 5;
  h.rng_seed = seed; h.rng_ctr = 0; h.ep_len = 0; h.episode = 0;
}

constexpr int OFF_RESET_SCRATCH = 116;  // shop part of the cold record, 52 bytes used
// apply the 51 parked Fisher-Yates draws in random.shuffle's order, then clear them
__device__ __forceinline__ void reset_blocks_finish(uint8_t* rec) {
#pragma unroll 1
  for (int i = 51; i >= 1; i--) {
    int j = rec[OFF_RESET_SCRATCH + i];
    int a = deck16(rec, i), b = deck16(rec, j);
    set_deck16(rec, i, b); set_deck16(rec, j, a);
  }
#pragma unroll 1
  for (int o = 112; o < BGYM_COLD_BYTES; o += 16) sts128(rec + o, make_uint4(0, 0, 0, 0));
}
// deck + shop blocks of a fresh episode, one lane doing all the work (reset kernel: every lane of
// the warp resets, so per-lane serial work is already converged).
//   deck52 != nullptr: replay of a supplied permutation (the reference's MT19937 shuffle stream)
//   else: suit-major build (balatro_env_2.py:519-522) + Fisher-Yates in random.shuffle's order
//         (for i in reversed(range(1, n)): j = randbelow(i + 1); swap) with native Philox draws
__device__ __noinline__ void reset_blocks_serial(uint8_t* rec, uint32_t seed, const uint8_t* deck52, bool gen) {
#pragma unroll 1
  for (int o = 0; o < BGYM_COLD_BYTES; o += 16) sts128(rec + o, make_uint4(0, 0, 0, 0));
  if (deck52) {
#pragma unroll 1
    for (int i = 0; i < 52; i++) {
      const int code = deck52[i];
      set_deck16(rec, i, code | (gen ? gen_card_mods(seed, (code & 3) * 13 + (code >> 2)) : 0));
    }
    return;
  }
  // ordered deck (+ generated modifiers, two cards per Philox block, two blocks per iteration), then all 51
  // Fisher-Yates draws (two blocks = four draws per iteration) parked as bytes in the still empty shop part of the
  // record, then the swaps in random.shuffle's order.  The blocks are independent: generating them in pairs, in line,
  // lets two dependency chains overlap (this loop is the latency of a level-2 reset tile).
#pragma unroll 1
  for (int k = 0; k < 52; k += 4) {
    int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    if (gen) {
      const uint4 p = philox4x32_10_inl((uint32_t)(k >> 1), 0, 0, 0, seed, BGYM_GEN_KEY1);
      const uint4 q = philox4x32_10_inl((uint32_t)(k >> 1) + 1, 0, 0, 0, seed, BGYM_GEN_KEY1);
      m0 = gen_mods_from_words(p.x, p.y); m1 = gen_mods_from_words(p.z, p.w);
      m2 = gen_mods_from_words(q.x, q.y); m3 = gen_mods_from_words(q.z, q.w);
    }
    set_deck16(rec, k, ((k % 13) * 4 + k / 13) | m0);
    set_deck16(rec, k + 1, (((k + 1) % 13) * 4 + (k + 1) / 13) | m1);
    set_deck16(rec, k + 2, (((k + 2) % 13) * 4 + (k + 2) / 13) | m2);
    set_deck16(rec, k + 3, (((k + 3) % 13) * 4 + (k + 3) / 13) | m3);
  }
#pragma unroll 1
  for (int b = 0; b < 26; b += 2) {       // block b -> draws for i = 2b+1, 2b+2
    const uint4 p = philox4x32_10_inl((uint32_t)b, 0, 0, 0, seed, BGYM_SHUFFLE_KEY1);
    const uint4 q = philox4x32_10_inl((uint32_t)b + 1, 0, 0, 0, seed, BGYM_SHUFFLE_KEY1);
    rec[OFF_RESET_SCRATCH + 2 * b + 1] = (uint8_t)shuffle_j_from_block(p, 2 * b + 1);
    rec[OFF_RESET_SCRATCH + 2 * b + 2] = (uint8_t)shuffle_j_from_block(p, 2 * b + 2);
    rec[OFF_RESET_SCRATCH + 2 * b + 3] = (uint8_t)shuffle_j_from_block(q, 2 * b + 3);
    if (2 * b + 4 <= 51) rec[OFF_RESET_SCRATCH + 2 * b + 4] = (uint8_t)shuffle_j_from_block(q, 2 * b + 4);
  }
  reset_blocks_finish(rec);
}

// In-kernel autoreset of the deck/shop blocks, split in two so that several terminated lanes of a
// warp cost little more than one:
//   prepare (whole warp, once per terminated env): zero the blocks, build the ordered deck and
//            compute all 51 Fisher-Yates draws lane-parallel (one Philox block per lane); the draws
//            are parked as bytes in the (just zeroed) shop block of the record;
//   finish  (each terminated lane for itself, all of them in parallel): apply the 51 swaps in
//            random.shuffle's order, then clear the parked draws.
__device__ __forceinline__ void reset_blocks_prepare(uint8_t* rec_of_src, uint32_t seed, int lane, bool gen) {
  if (lane < 11) sts128(rec_of_src + 16 * lane, make_uint4(0, 0, 0, 0));
  __syncwarp();
  // generated modifiers are attached to the card before the shuffle and travel with it (include/bgym.h)
  set_deck16(rec_of_src, lane, ((lane % 13) * 4 + lane / 13) | (gen ? gen_card_mods(seed, lane) : 0));
  if (lane < 20) set_deck16(rec_of_src, lane + 32, (((lane + 32) % 13) * 4 + (lane + 32) / 13) | (gen ? gen_card_mods(seed, lane + 32) : 0));
  // lane l (< 26) owns block l -> draws for i = 2l+1 and i = 2l+2
  uint4 blk = philox4x32_10((uint32_t)lane, 0, 0, 0, seed, BGYM_SHUFFLE_KEY1);
  if (lane < 26) {
    rec_of_src[OFF_RESET_SCRATCH + 2 * lane + 1] = (uint8_t)shuffle_j_from_block(blk, 2 * lane + 1);
    if (2 * lane + 2 <= 51) rec_of_src[OFF_RESET_SCRATCH + 2 * lane + 2] = (uint8_t)shuffle_j_from_block(blk, 2 * lane + 2);
  }
}
// warp-level driver: `want_reset` lanes get fresh deck/shop blocks in their record `my_rec`
__device__ __forceinline__ void autoreset_warp(bool want_reset, uint32_t new_seed, uint8_t* my_rec, int lane, bool gen) {
  uint32_t rmask = __ballot_sync(0xffffffffu, want_reset);
  if (!rmask) return;
  __syncwarp();     // a lane's own writes to its record (the step it just ran) are ordered before the warp rewrites that record
  unsigned long long my_ptr = reinterpret_cast<unsigned long long>(my_rec);
  uint32_t m = rmask;
  while (m) {
    int src = __ffs(m) - 1;
    m &= m - 1;
    uint32_t sd = __shfl_sync(0xffffffffu, new_seed, src);
    uint8_t* rec_src = reinterpret_cast<uint8_t*>(__shfl_sync(0xffffffffu, my_ptr, src));
    reset_blocks_prepare(rec_src, sd, lane, gen);
  }
  __syncwarp();
  if (want_reset) reset_blocks_finish(my_rec);
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// rare paths, out of line.  They work on the PACKED record so the step kernel's hot instruction
// stream stays small and register-resident: pack_hot -> rare_dispatch -> unpack_hot at one site.
// ---------------------------------------------------------------------------------------------
enum { RARE_NONE = 0, RARE_ADVANCE = 1, RARE_REROLL = 2, RARE_CONSUMABLE = 3 };
struct RareOut { double reward; int err; int terminated; uint64_t immolate_removed; };

__device__ __noinline__ void rare_dispatch(uint8_t* hot, uint8_t* rec, int op, int arg, Draws* rng, RareOut* out) {
  Hot h;
  unpack_hot(hot, h);
  out->reward = 0.0; out->err = 0; out->terminated = 0; out->immolate_removed = 0;
  if (op == RARE_ADVANCE) {
    advance_round(h, rec, *rng);
  } else if (op == RARE_REROLL) {
    int* rr = reinterpret_cast<int*>(rec + OFF_REROLL);
    int cost = (int)(*rr * shop_cost_mult(h));  // shop.py:172
    if (h.money < cost) { out->reward = -1.0; out->err = BGYM_ERR_SHOP; }
    else {
      h.money -= cost;
      *rr = (int)(*rr * 1.35);
      shop_generate_inventory(h, rec, *rng);
    }
  } else if (op == RARE_CONSUMABLE) {
    // (Immolate moves cards under the hand's deck indices: the tile refreshes the hand codes after immolate_compact)
    out->reward = use_consumable(h, hot, rec, arg, *rng, out->err, out->terminated, &out->immolate_removed);
  }
  pack_hot(hot, h);
}

// the ante > 3 score reward goes through log10 (balatro_env_2.py:821); kept out of line
__device__ __noinline__ double log10_out_of_line(double x) { return log10(x); }

// ---------------------------------------------------------------------------------------------
// per-step outputs
// ---------------------------------------------------------------------------------------------
struct StepInfo {
  long long final_score;
  double x_mult;
  int chips, mult, hand_type, error_code, flags, cards_played, base_score;
};

// ---------------------------------------------------------------------------------------------
// the step, balatro_env_2.py:616-1064 (play), :1174-1253 (shop), :1255-1318 (blind select)
// ---------------------------------------------------------------------------------------------
// CATS: which action categories this instantiation compiles in (the partitioned step runs one small
// kernel per category so that every warp in flight executes the same short code).
//   CAT_CONS = USE_CONSUMABLE, CAT_SHOP = buy / sell / end shop, CAT_BLIND = select blind, CAT_GEN = reroll / skip blind
//   (the two actions that generate a shop inventory).
// DEFER_ADVANCE: a played hand that beats the blind stops before the round advance (advance_round + shop generation):
//   *defer_snap receives the draw-sequence position (Draws::snapshot) and h.rng_ctr the raw block counter; the caller hands
//   the env to a second-level tile that runs step_env_advance() — there every lane of the warp advances a round, whereas
//   inside the PLAY tile a few lanes would walk that long path while the others wait.  -1 = nothing deferred.
//   CAT_SHOP_END = leaving the shop: like a blind selection it deals a hand (draw + hand codes), so it rides with CAT_BLIND
enum { CAT_SELECT = 1, CAT_PLAY = 2, CAT_DISCARD = 4, CAT_CONS = 8, CAT_SHOP = 16, CAT_BLIND = 32, CAT_GEN = 64, CAT_SHOP_END = 128,
       CAT_OTHER = CAT_CONS | CAT_SHOP | CAT_BLIND | CAT_GEN | CAT_SHOP_END, CAT_ALL = 255 };
template <int CATS, bool DEFER_ADVANCE>
__device__ void step_env(Hot& h, uint8_t* hot, uint8_t* rec, int action, uint64_t mask, const BgymDraws* tape, double& reward_out,
                         int& terminated_out, StepInfo& info, int* defer_snap, uint64_t* immolate_removed = nullptr) {
  if (DEFER_ADVANCE) *defer_snap = -1;
  info.final_score = 0; info.x_mult = 1.0; info.chips = 0; info.mult = 0; info.hand_type = -1;
  info.error_code = 0; info.flags = 0; info.cards_played = 0; info.base_score = 0;
  reward_out = 0.0; terminated_out = 0;
  if (h.ante > 100 || h.chips_scored > 1000000000LL) {  // :619-623
    info.flags = BGYM_F_GUARD_TERMINATED; terminated_out = 1;
    return;
  }
  if (action < 0 || action >= BGYM_NUM_ACTIONS || !((mask >> action) & 1)) {  // :626-627
    info.error_code = BGYM_ERR_INVALID_ACTION; reward_out = -1.0;
    return;
  }
  h.ep_len++;
  Draws rng;
  rng.init(h.rng_seed, h.rng_ctr, tape);
  // glass / lucky rolls of a played hand's card loop, the draws of a consumable: ONE converged Philox call here instead
  // of one per lane inside whichever branch draws first (an unused block is not counted, Draws::blocks)
  if (CATS & (CAT_PLAY | CAT_CONS)) rng.prefetch();
  double reward = 0.0;
  int terminated = 0;
  int rare_op = RARE_NONE, rare_arg = 0;
  bool hand_changed = false;

  if ((CATS & CAT_SELECT) && action >= BGYM_A_SELECT_BASE && action < BGYM_A_SELECT_BASE + 8) {
    // toggle in the ordered selection list (:1052-1058); legal only in PLAY phase by the mask
    int slot = action - BGYM_A_SELECT_BASE;
    const int found = sel_find(h.sel_order, h.sel_n, slot);
    if (found >= 0) {
      uint32_t lo = h.sel_order & ((1u << (4 * found)) - 1);
      uint32_t hi = (found == 7) ? 0u : ((h.sel_order >> (4 * (found + 1))) << (4 * found));
      h.sel_order = lo | hi; h.sel_n--;
    } else {
      h.sel_order |= (uint32_t)slot << (4 * h.sel_n); h.sel_n++;
    }
  } else if ((CATS & CAT_PLAY) && action == BGYM_A_PLAY_HAND) {
    // ---- gather the played cards in selection order (:650-660) ----
    int sel_slots = 0;                 // bit set of selected hand slots
    uint64_t played_bits = 0;          // bit set over deck indices of the played cards
    int n_played = 0, card_chip_sum = 0, faces_ge11 = 0, extra_money = 0, n_red = 0, n_blue = 0;
    int boss = h.boss_type, debuffed = 0;
    int n_rolls = 0;            // u01 draws of the card loop, in order; bit j of lucky_rolls: draw j decides a Lucky card's money
    uint32_t lucky_rolls = 0;
    #pragma unroll 1
    for (int k = 0; k < h.sel_n; k++) {
      int sl = nib_at(h.sel_order, k);
      if (sl >= h.hand_n) continue;
      sel_slots |= 1 << sl;
      int idx = byte_at(h.hand, sl);
      if (idx >= h.deck_n) continue;
      int c = deck16(rec, idx);
      int code = c16_code(c), enh = c16_enh(c), seal = c16_seal(c);
      n_played++;
      played_bits |= 1ull << idx;
      card_chip_sum += card_chips(code, enh, c16_edition(c));
      int r = code >> 2;                       // rank - 2
      faces_ge11 += r >= 9;                    // J Q K A (:862 counts rank.value >= 11)
      bool jqk = r >= 9 && r <= 11;
      // boss debuffs (boss_blinds.py:447-478): Plant = face ranks, Violet = all, Pillar = played before
      debuffed += (boss == B_PLANT && jqk) || boss == B_VIOLET || (boss == B_PILLAR && ((h.boss_played_cards >> idx) & 1));
      // per-card enhancement / seal loop (:703-734), draws in selection order
      // (:703-734 draws one uniform for a Glass card and two for a Lucky card, of which only the second decides
      // anything; the draws themselves are made after the loop, by all lanes that have some at once)
      if (enh == BGYM_ENH_GLASS) n_rolls++;
      else if (enh == BGYM_ENH_LUCKY) { lucky_rolls |= 2u << n_rolls; n_rolls += 2; }
      extra_money += (seal == BGYM_SEAL_GOLD) ? 3 : 0;
      n_red += seal == BGYM_SEAL_RED;
      // blue seal: room is tested against the list as it is now, creation re-tests (:733, :765-767)
      n_blue += (seal == BGYM_SEAL_BLUE) && (h.cons_n < h.cons_slots);
    }
    if (n_rolls) extra_money += 20 * rng.count_u01_below(n_rolls, lucky_rolls, 0.0667);
    // ---- highlight + classification on deck[slot] of every highlighted slot (:663-671) ----
    h.highlight |= sel_slots;
    HandHist hist;
    hist.clear();
    {
      const uint4 d8 = *reinterpret_cast<const uint4*>(rec + OFF_DECK);      // deck[0..7], one load
      const uint32_t dw[4] = {d8.x, d8.y, d8.z, d8.w};
      #pragma unroll
      for (int sl = 0; sl < 8; sl++)
        if ((h.highlight >> sl) & 1) hist.add(c16_code((int)((dw[sl >> 1] >> (16 * (sl & 1))) & 0xFFFFu)));
    }
    int ht = classify(hist);
    // ---- boss gate (boss_blinds.py:380-407) ----
    bool allowed = true;
    if (boss == B_PSYCHIC) allowed = n_played == 5;
    else if (boss == B_EYE) allowed = !((h.boss_played_types >> ht) & 1);
    else if (boss == B_MOUTH) allowed = (h.boss_played_types == 0) || ((h.boss_played_types >> ht) & 1);
    else if (boss == B_VERDANT) allowed = n_played >= h.boss_cards_required;
    if (!allowed) {
      info.error_code = BGYM_ERR_BOSS_RESTRICTION; reward_out = -1.0;
      return;  // highlight stays modified, exactly as in the reference
    }
    // ---- score: jokers never fire in-env, x_mult = 1.0 ----
    int bc, bm;
    hand_base(ht, hand_level(h, ht), bc, bm);
    int chips = bc + card_chip_sum, mult = bm;
    long long base_score = (long long)chips * mult;
    long long fs = base_score;
    // steel cards left in hand (:560-570)
    int n_steel = 0;
    #pragma unroll
    for (int i = 0; i < 8; i++) {
      const int idx = (int)((h.hand >> (8 * i)) & 0xFF);
      if (i < h.hand_n && !((sel_slots >> i) & 1) && idx < 52 && c16_enh(deck16(rec, idx)) == BGYM_ENH_STEEL) n_steel++;
    }
    fs = (long long)((double)fs * c_pow_1_5[n_steel]);
    // boss modification ratio (:745-755, boss_blinds.py:409-445)
    if (boss) {
      int mc = bc, mm = bm;
      if (boss == B_FLINT) { mc = mc / 2; mm = mm / 2; }
      else if (boss == B_OXIDE) { mc = 0; }
      else if (boss == B_ARM) { mc = (int)(mc * 0.75); mm = (int)(mm * 0.75); }
      if (debuffed > 0) { double p = c_pow_0_8[debuffed]; mc = (int)(mc * p); mm = (int)(mm * p); }
      double cr = div_nz((double)mc, (double)bc), mr = div_nz((double)mm, (double)bm);
      fs = (long long)((double)fs * cr * mr);
    }
    fs = (long long)((double)fs * (1 + n_red * 0.5));  // retriggers :757-759
    h.money += extra_money;
    {  // blue seals -> planet of this hand type (cards.py:228-246)
      int planet = ht == 0 ? 38 : (ht <= 8 ? 29 + ht : 30 + ht);
      #pragma unroll 1
      for (int i = 0; i < n_blue; i++) if (h.cons_n < h.cons_slots) cons_append(h, planet);
    }
    double needed = (double)max(h.chips_needed, 1);
    double old_progress = fmin(1.0, div_nz((double)h.round_chips, needed));
    h.round_chips += fs; h.chips_scored += fs;
    h.hands_played_total++; h.hands_played_ante++;
    if (fs > (long long)h.best_hand) h.best_hand = (int)min(fs, 2147483647LL);
    if (rec[OFF_HPC + ht] < 255) rec[OFF_HPC + ht]++;
    if (boss) {  // on_hand_scored boss_blinds.py:480-507
      h.boss_played_types |= 1 << ht;
      h.boss_flags &= ~1;
      h.boss_hands_played = (h.boss_hands_played + 1) & 0xFF;
      if (boss == B_PILLAR) h.boss_played_cards |= played_bits;
      if (boss == B_VERDANT) h.boss_cards_required = min(7, h.boss_cards_required + 1);
    }
    h.sel_n = 0; h.sel_order = 0;
    // ---- reward shaping (:799-892) ----
    double new_progress = fmin(1.0, div_nz((double)h.round_chips, needed));
    double milestone = 0.0;
    if (old_progress < 0.25 && 0.25 <= new_progress) milestone = 5.0;
    else if (old_progress < 0.5 && 0.5 <= new_progress) milestone = 10.0;
    else if (old_progress < 0.75 && 0.75 <= new_progress) milestone = 15.0;
    else if (old_progress < 1.0 && 1.0 <= new_progress) milestone = 25.0;
    double score_reward = (h.ante <= 3) ? fmin(10.0, div_nz((double)fs, 100.0))
                                        : fmin(10.0, 3.0 * log10_out_of_line((double)max(fs, 1LL)));
    double hq = ht == 0 ? 0.1 : ht == 1 ? 0.5 : ht == 2 ? 1.0 : ht == 3 ? 2.0 : (ht == 4 || ht == 5) ? 2.5
              : ht == 6 ? 3.5 : ht == 7 ? 5.0 : ht == 8 ? 7.0 : ht == 9 ? 10.0 : 0.0;
    double eff = 0.0;
    if (ht >= BGYM_HT_THREE_KIND && n_played <= 3) eff = 2.0;
    else if (ht >= BGYM_HT_FLUSH && n_played == 5) eff = 1.0;
    else if (n_played <= 4 && h.hands_left <= 2) eff = 1.5;
    double syn = 0.0;
    if (h.joker_n > 0) {
      if (ht == BGYM_HT_FLUSH && (owns_joker(h, BGYM_J_SMEARED_JOKER) || owns_joker(h, BGYM_J_FOUR_FINGERS) || owns_joker(h, BGYM_J_SHORTCUT))) syn += 2.0;
      if (ht >= BGYM_HT_ONE_PAIR && ht <= BGYM_HT_THREE_KIND &&
          (owns_joker(h, BGYM_J_ODD_TODD) || owns_joker(h, BGYM_J_EVEN_STEVEN) || owns_joker(h, BGYM_J_JOLLY_JOKER) || owns_joker(h, BGYM_J_ZANY_JOKER))) syn += 1.5;
      if (faces_ge11 > 0 && (owns_joker(h, BGYM_J_SCARY_FACE) || owns_joker(h, BGYM_J_SMILEY_FACE) || owns_joker(h, BGYM_J_BUSINESS_CARD))) syn += 0.5 * faces_ge11;
    }
    double strat = 0.0;
    if (new_progress > 0.7 && h.hands_left >= 3) strat = 2.0;
    else if (new_progress < 0.3 && ht >= BGYM_HT_FLUSH) strat = 3.0;
    double ante_bonus = (h.ante >= 4) ? fmin(5.0, (h.ante - 3) * 0.5) : 0.0;
    reward = 15.0 * new_progress + milestone + score_reward + hq * 2.0 + eff * 1.5 + syn * 3.0 + strat * 2.0 + ante_bonus;
    reward = fmin(reward, 100.0);
    info.final_score = fs; info.base_score = (int)min(base_score, 2147483647LL); info.chips = chips; info.mult = mult;
    info.hand_type = ht; info.cards_played = n_played; info.flags |= BGYM_F_PLAYED;
    // ---- round end (:914-960) ----
    if (h.round_chips >= (long long)h.chips_needed) {
      reward += fmin(50.0, 25.0 + 10.0 * h.ante);
      rare_op = RARE_ADVANCE;
      info.flags |= BGYM_F_BEAT_BLIND;
    } else if (h.hands_left <= 1) {
      reward += -50.0 * (1.0 - new_progress);
      terminated = 1;
      info.flags |= BGYM_F_FAILED;
    } else {
      h.hands_left--;
      draw_cards(h);
      hand_changed = true;
      if (boss) {  // on_hand_drawn boss_blinds.py:343-378 (first_hand is already False here)
        int face = 0;
        if (boss == B_HOOK) {
          if (h.hand_n >= 2) {
            int a, b;
            if (rng.tape) { a = rng.below(h.hand_n); b = rng.below(h.hand_n); }
            else { a = rng.below(h.hand_n); b = rng.below(h.hand_n - 1); b += (b >= a); }
            remove_slots(h, (1 << a) | (1 << b));
          }
        } else if (boss == B_WHEEL) {
          #pragma unroll 1
          for (int i = 0; i < h.hand_n; i++) if (rng.u01() < 1.0 / 7) face |= 1 << i;
        } else if (boss == B_MARK) {
          #pragma unroll 1
          for (int i = 0; i < h.hand_n; i++) {
            int r = c16_code(deck16(rec, byte_at(h.hand, i))) >> 2;
            if (r >= 9 && r <= 11) face |= 1 << i;
          }
        } else if (boss == B_FISH) {
          face = (1 << h.hand_n) - 1;
        }
        h.face_down = face;
      }
    }
  } else if ((CATS & CAT_DISCARD) && action == BGYM_A_DISCARD) {
    // ---- :962-1050 ----
    int sel_slots = 0, n_disc = 0, purple = 0, faces = 0;
    #pragma unroll 1
    for (int k = 0; k < h.sel_n; k++) {
      int sl = nib_at(h.sel_order, k);
      if (sl >= h.hand_n) continue;
      sel_slots |= 1 << sl;
      int idx = byte_at(h.hand, sl);
      if (idx >= h.deck_n) continue;
      int c = deck16(rec, idx);
      int r = c16_code(c) >> 2;
      n_disc++;
      purple += c16_seal(c) == BGYM_SEAL_PURPLE;
      faces += r >= 9 && r <= 11;
    }
    int money_from_discards = 0, n_discard_jokers = 0;
    #pragma unroll 1
    for (int j = 0; j < h.joker_n; j++) {  // discard-phase joker effects, complete_joker_effects.py:186-209
      int id = byte_at(h.jokers, j);
      int money = (id == BGYM_J_TRADING_CARD && h.discards_left == 3 && n_disc == 1) ? 3
                : (id == BGYM_J_FACELESS_JOKER && faces >= 3) ? 5 : 0;
      money_from_discards += money;
      n_discard_jokers += (id == BGYM_J_FACELESS_JOKER || id == BGYM_J_HIT_THE_ROAD || id == BGYM_J_RESERVED_PARKING || id == BGYM_J_LUCHADOR);
    }
    h.money += money_from_discards;
    // discard_hand (balatro_game.py:111-127): every highlighted slot goes, stale ones too
    remove_slots(h, h.highlight | sel_slots);
    h.highlight = 0;
    draw_cards(h);
    hand_changed = true;
    h.discards_left--;
    h.sel_n = 0; h.sel_order = 0;
    #pragma unroll 1
    for (int i = 0; i < purple; i++)  // purple seals -> tarots (:1021-1032)
      if (h.cons_n < h.cons_slots) cons_append(h, BGYM_CONS_TAROT_BASE + rng.below(22));
    reward = 0.2;
    if (n_discard_jokers) reward += 0.5 * n_discard_jokers;
    if (money_from_discards > 0) reward += money_from_discards / 5.0;
    double progress = div_nz((double)h.round_chips, (double)max(h.chips_needed, 1));
    if (progress < 0.5 && h.discards_left > 1) reward += 0.5;
    else if (progress > 0.8 && h.discards_left > 1) reward -= 0.3;
  } else if ((CATS & CAT_CONS) && action >= BGYM_A_USE_CONS_BASE && action < BGYM_A_USE_CONS_BASE + 5) {
    rare_op = RARE_CONSUMABLE; rare_arg = action - BGYM_A_USE_CONS_BASE;
  } else if ((CATS & CAT_SHOP_END) && action == BGYM_A_SHOP_END) {
    h.phase = BGYM_PHASE_PLAY;
    draw_cards(h);
    hand_changed = true;
    info.flags |= BGYM_F_SHOP_DONE;
  } else if ((CATS & CAT_GEN) && action == BGYM_A_SHOP_REROLL) {
    rare_op = RARE_REROLL;
  } else if ((CATS & CAT_SHOP) && action >= BGYM_A_SHOP_BUY_BASE && action < BGYM_A_SHOP_BUY_BASE + 10) {
    int i = action - BGYM_A_SHOP_BUY_BASE;
    int n_items = rec[OFF_N_ITEMS];
    int type = rec[OFF_ITEM_TYPE + i], id = rec[OFF_ITEM_ID + i];
    int cost = *reinterpret_cast<int*>(rec + OFF_ITEM_COST + 4 * i);
    h.money -= cost;  // shop.py:185-187: pay, pop the item
    #pragma unroll 1
    for (int k = i; k + 1 < n_items; k++)
      shop_put(rec, k, rec[OFF_ITEM_TYPE + k + 1], rec[OFF_ITEM_ID + k + 1], *reinterpret_cast<int*>(rec + OFF_ITEM_COST + 4 * (k + 1)));
    shop_put(rec, n_items - 1, 0, 0, 0);
    rec[OFF_N_ITEMS] = (uint8_t)(n_items - 1);
    if (type == BGYM_ITEM_PACK) {
      int count = (id == BGYM_PACK_STANDARD) ? 3 : 1;  // shop.py:150-157: cards go to player.deck only
      #pragma unroll 1
      for (int k = 0; k < count; k++) (void)rng.below(52);
      reward = 5.0;
    } else if (type == BGYM_ITEM_CARD) {
      reward = 3.0;
    } else if (type == BGYM_ITEM_JOKER) {
      if (h.joker_n >= 5) { reward = -1.0; info.error_code = BGYM_ERR_SHOP; }  // shop.py:196-197
      else { h.jokers = with_byte(h.jokers, h.joker_n, id); h.joker_n++; reward = 15.0; }
    } else {
      if (id == BGYM_VOUCHER_MAGIC_TRICK) h.n_magic = min(h.n_magic + 1, 255); else h.n_minimalist = min(h.n_minimalist + 1, 255);
      reward = 10.0;
    }
  } else if ((CATS & CAT_SHOP) && action >= BGYM_A_SELL_JOKER_BASE && action < BGYM_A_SELL_JOKER_BASE + 5) {
    int j = action - BGYM_A_SELL_JOKER_BASE;
    int id = byte_at(h.jokers, j);
    uint64_t lo = h.jokers & ((1ull << (8 * j)) - 1);
    uint64_t hi = (j >= 7) ? 0 : (h.jokers >> (8 * (j + 1))) << (8 * j);
    h.jokers = lo | hi; h.joker_n--;
    int sell = max(3, (int)c_joker_cost[id] / 2);
    h.money += sell; h.jokers_sold++;
    reward = sell / 5.0;
  } else if ((CATS & CAT_BLIND) && action >= BGYM_A_SELECT_BLIND_BASE && action < BGYM_A_SELECT_BLIND_BASE + 3) {
    int bt = action - BGYM_A_SELECT_BLIND_BASE;
    h.round = bt + 1;
    long long needed = (h.ante <= 8) ? (long long)c_blind_chips[max(h.ante, 1) - 1][bt]
                                     : (long long)(c_blind_chips[7][bt] * c_pow_1_5[min(h.ante - 8, 100)]);
    if (bt == 2) {
      int boss = 1 + rng.below(28);  // select_boss_blind boss_blinds.py:522-532
      h.boss_type = boss;            // activate_boss_blind :308-341
      h.boss_flags = 1; h.boss_cards_required = 5; h.boss_played_types = 0; h.boss_hands_played = 0; h.boss_played_cards = 0;
      needed = (long long)((double)needed * (boss == B_WALL ? 2.0 : 1.0));
      if (boss == B_WATER) h.discards_left = 0;
      if (boss == B_MANACLE) h.hand_size -= 1;
      if (boss == B_NEEDLE) h.hands_left = 1;
      reward = 10.0;
    }
    h.chips_needed = (int)min(needed, 2147483647LL);
    h.phase = BGYM_PHASE_PLAY;
    draw_cards(h);
    hand_changed = true;
  } else if ((CATS & CAT_GEN) && action == BGYM_A_SKIP_BLIND) {
    reward = -5.0;
    rare_op = RARE_ADVANCE;
  }
  if (DEFER_ADVANCE && (CATS & CAT_PLAY) && rare_op == RARE_ADVANCE && action == BGYM_A_PLAY_HAND) {
    *defer_snap = (int)rng.snapshot();
    if (!tape) h.rng_ctr = rng.ctr;          // raw counter: step_env_advance() settles it (Draws::blocks) when it is done
    reward_out = reward;
    terminated_out = terminated;
    return;                                  // a beaten blind changes neither the hand nor the selection any further
  }
  if ((CATS & (CAT_PLAY | CAT_CONS | CAT_GEN)) && rare_op != RARE_NONE) {   // ONE out-of-line site for every rare path
    RareOut ro;
    pack_hot(hot, h);
    rare_dispatch(hot, rec, rare_op, rare_arg, &rng, &ro);
    unpack_hot(hot, h);
    if (rare_op != RARE_ADVANCE) { reward = ro.reward; info.error_code = ro.err; terminated = ro.terminated; }
    if (CATS & CAT_CONS) *immolate_removed = ro.immolate_removed;
  }
  if ((CATS & (CAT_PLAY | CAT_DISCARD | CAT_OTHER)) && hand_changed) refresh_hand_codes(h, rec);
  if (!tape) h.rng_ctr = rng.blocks();
  reward_out = reward;
  terminated_out = terminated;
}

// second half of a step whose round advance was deferred (DEFER_ADVANCE above): balatro_env_2.py:914-925 -> :1326-1392
__device__ __forceinline__ void step_env_advance(Hot& h, uint8_t* hot, uint8_t* rec, const BgymDraws* tape, uint32_t snap) {
  Draws rng;
  rng.restore(h.rng_seed, h.rng_ctr, tape, snap);
  RareOut ro;
  pack_hot(hot, h);
  rare_dispatch(hot, rec, RARE_ADVANCE, 0, &rng, &ro);
  unpack_hot(hot, h);
  if (!tape) h.rng_ctr = rng.blocks();
}

// ---------------------------------------------------------------------------------------------
// observation record (balatro_env_2.py:1473-1573) assembled in registers, 11 x 16 B
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int obs_cons_id(int cid) { return cid >= BGYM_CONS_ENUMSTYLE_BASE ? 0 : cid; }

// the shop block of the observation (shop_items[10], shop_costs[10] as int16 pairs in record order):
// the only part of the observation that reads the cold record
struct ShopObs { uint32_t w[11]; };   // it0 | it1,2 | it3,4 | it5,6 | it7,8 | it9,ic0 | ic1,2 | ic3,4 | ic5,6 | ic7,8 | ic9

__device__ __forceinline__ void obs_shop_block(const Hot& h, const uint8_t* rec, ShopObs& so) {
  if (h.phase != BGYM_PHASE_SHOP) {       // nothing to read: a whole tile of the PLAY-phase lists takes this branch
#pragma unroll
    for (int i = 0; i < 11; i++) so.w[i] = 0;
    return;
  }
  int n_items = rec[OFF_N_ITEMS];
  int it[10], ic[10];
#pragma unroll
  for (int i = 0; i < 10; i++) {
    bool on = i < n_items && i < 9;
    it[i] = on ? rec[OFF_ITEM_TYPE + (i < 9 ? i : 8)] : 0;
    ic[i] = on ? (*reinterpret_cast<const int*>(rec + OFF_ITEM_COST + 4 * (i < 9 ? i : 8)) & 0xFFFF) : 0;
  }
  so.w[0] = it[0];
  so.w[1] = it[1] | (it[2] << 16); so.w[2] = it[3] | (it[4] << 16); so.w[3] = it[5] | (it[6] << 16); so.w[4] = it[7] | (it[8] << 16);
  so.w[5] = it[9] | (ic[0] << 16);
  so.w[6] = ic[1] | (ic[2] << 16); so.w[7] = ic[3] | (ic[4] << 16); so.w[8] = ic[5] | (ic[6] << 16);
  so.w[9] = ic[7] | (ic[8] << 16); so.w[10] = ic[9];
}

template <bool CG = false>
__device__ __forceinline__ void write_obs_regs(const Hot& h, const ShopObs& so, uint64_t mask, uint8_t* obs /*smem, 176 B*/) {
  uint4 q;
  const uint32_t selm = sel_slot_mask(h.sel_order, h.sel_n);
  // 0: hand[8] | selected_cards[8]   (hand codes = deck[hand[i]], -1 when empty)
  q.x = (uint32_t)h.hand_code; q.y = (uint32_t)(h.hand_code >> 32);
  q.z = spread4(selm); q.w = spread4(selm >> 4);
  str128<CG>(obs, q);
  // 16: face_down_cards[8] | chips_scored
  q.x = spread4(h.face_down); q.y = spread4(h.face_down >> 4);
  q.z = (uint32_t)h.chips_scored; q.w = (uint32_t)((uint64_t)h.chips_scored >> 32);
  str128<CG>(obs + 16, q);
  // 32: round_chips_scored, progress_ratio, mult, chips_needed
  double p = div_nz((double)h.round_chips, (double)max(h.chips_needed, 1));
  q.x = (uint32_t)h.round_chips; q.y = __float_as_uint((float)fmin(p, 2.0)); q.z = 1u; q.w = (uint32_t)h.chips_needed;
  str128<CG>(obs + 32, q);
  // 48: money, hands_played, best_hand_this_ante, ante | shop_rerolls
  q.x = (uint32_t)h.money; q.y = (uint32_t)h.hands_played_total; q.z = (uint32_t)h.best_hand;
  q.w = (h.ante & 0xFFFF) | ((uint32_t)(h.shop_reroll_state & 0xFFFF) << 16);
  str128<CG>(obs + 48, q);
  // 64..159: int16 arrays and int8 scalars, built as 16-bit lanes
  // joker_ids[10] @64
  uint32_t j[5];
#pragma unroll
  for (int i = 0; i < 4; i++) j[i] = (uint32_t)byte_at(h.jokers, 2 * i) | ((uint32_t)byte_at(h.jokers, 2 * i + 1) << 16);
  q.x = j[0]; q.y = j[1]; q.z = j[2]; q.w = j[3];
  str128<CG>(obs + 64, q);
  // 80: joker_ids[8..9] (always 0) | consumables[0..4] (84..93) | shop_items[0] (94)
  int c0 = h.cons_n > 0 ? obs_cons_id(byte_at(h.cons, 0)) : 0, c1 = h.cons_n > 1 ? obs_cons_id(byte_at(h.cons, 1)) : 0;
  int c2 = h.cons_n > 2 ? obs_cons_id(byte_at(h.cons, 2)) : 0, c3 = h.cons_n > 3 ? obs_cons_id(byte_at(h.cons, 3)) : 0;
  int c4 = h.cons_n > 4 ? obs_cons_id(byte_at(h.cons, 4)) : 0;
  q.x = 0; q.y = (uint32_t)c0 | ((uint32_t)c1 << 16); q.z = (uint32_t)c2 | ((uint32_t)c3 << 16); q.w = (uint32_t)c4 | (so.w[0] << 16);
  str128<CG>(obs + 80, q);
  // 96: shop_items[1..8]
  q.x = so.w[1]; q.y = so.w[2]; q.z = so.w[3]; q.w = so.w[4];
  str128<CG>(obs + 96, q);
  // 112: shop_items[9] | shop_costs[0..6]
  q.x = so.w[5]; q.y = so.w[6]; q.z = so.w[7]; q.w = so.w[8];
  str128<CG>(obs + 112, q);
  // 128: shop_costs[7..9] (128..133) | hand_levels[0..9] (134..143)
  q.x = so.w[9];
  q.y = so.w[10] | ((h.lv0 & 0xFFFF) << 16);
  q.z = (h.lv0 >> 16) | ((h.lv1 & 0xFFFF) << 16);
  q.w = (h.lv1 >> 16) | ((h.lv2 & 0xFFFF) << 16);
  str128<CG>(obs + 128, q);
  // 144: hand_levels[10..11] | hand_size, deck_size | round, hands_left, discards_left, joker_count |
  //      joker_slots, consumable_count, consumable_slots, phase | boss_active, boss_type, pad, pad
  q.x = (h.lv2 >> 16) | ((uint32_t)(h.hand_n & 0xFF) << 16) | ((uint32_t)(h.deck_n & 0xFF) << 24);
  q.y = (h.round & 0xFF) | ((h.hands_left & 0xFF) << 8) | ((h.discards_left & 0xFF) << 16) | ((uint32_t)(h.joker_n & 0xFF) << 24);
  q.z = (h.joker_slots & 0xFF) | ((h.cons_n & 0xFF) << 8) | ((h.cons_slots & 0xFF) << 16) | ((uint32_t)(h.phase & 0xFF) << 24);
  q.w = (h.boss_type != 0 ? 1u : 0u) | ((uint32_t)(h.boss_type & 0xFF) << 8);
  str128<CG>(obs + 144, q);
  // 160: action_mask_bits | pad
  q.x = (uint32_t)mask; q.y = (uint32_t)(mask >> 32); q.z = 0; q.w = 0;
  str128<CG>(obs + 160, q);
}

// selection record (BgymSel, 16 B): selected_cards[8] | action_mask_bits — the two observation fields a toggle changes
__device__ __forceinline__ uint4 sel_words(const Hot& h, uint64_t mask) {
  const uint32_t selm = sel_slot_mask(h.sel_order, h.sel_n);
  uint4 q;
  q.x = spread4(selm); q.y = spread4(selm >> 4); q.z = (uint32_t)mask; q.w = (uint32_t)(mask >> 32);
  return q;
}

template <bool CG = false>
__device__ __forceinline__ void write_obs(const Hot& h, const uint8_t* rec, uint64_t mask, uint8_t* obs) {
  ShopObs so;
  if (rec) obs_shop_block(h, rec, so);
  else {
#pragma unroll
    for (int i = 0; i < 11; i++) so.w[i] = 0;   // PLAY phase (main pass): the shop block is all zeros
  }
  write_obs_regs<CG>(h, so, mask, obs);
}

}  // namespace bgym
