// bgym_kernels.cu — sm_100a kernels and the C-ABI (include/bgym.h) of the batched Balatro env.
//
// Kernels
//   env_step_main_kernel / env_step_list_kernel<LIST> / env_step_small_kernel   K2  the fused, path-partitioned BalatroEnv.step
//                           (bgym_step_part.cuh): mask check -> phase dispatch -> scoring -> boss ->
//                           round advance / shop generation -> reward -> observation + mask emission,
//                           in-place autoreset (K3) and an optional fused random-legal policy
//   env_reset_kernel    K3  reset + first observation
//   score_hands5_kernel / score_hands_kernel   K1  classify + chips x mult + joker interpreter (K4)
//   action_mask_kernel, sample_actions_kernel, episode_stats_kernel (K6)
//   featurize_kernel, masked_sample_kernel, gae_kernel (bgym_rollout.cuh)   on-device PPO rollout collection
//
//   sync_state / sync_obs kernels   fold the device-only side arrays (BgymTog, BgymSel) into the records and back
//   pack_count / pack_index / pack_records / scatter_dirty kernels   observation deltas for a host mirror (HostMirror)
//
// Data layout of K2/K3 (include/bgym.h): env state = hot[n] (144 B records), cold[n] (176 B) and tog[n] (32 B toggle
// records: the card-select working set); observations = obs[n] (176 B) and sel[n] (16 B selection records).  A card
// toggle — three steps in four — reads and writes tog and sel only, coalesced, from registers (main pass).  Every other
// action goes through a list tile: one lane per env, records read and written where they lie.  The reset kernel moves
// tiles of records with 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) completing on a per-warp mbarrier; record
// strides 144 / 176 B are odd multiples of 16 B, so per-lane 128-bit shared-memory accesses are bank-conflict free.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "bgym_env.cuh"

namespace bgym {

struct StepArgs {
  uint8_t* hot;              // n x 144
  uint8_t* tog;              // n x 32  toggle records (BgymTog): the select path's working set
  uint8_t* cold;             // n x 176
  const int32_t* actions;    // n (read; written instead with BGYM_FLAG_RANDOM_POLICY)
  int32_t* actions_out;      // n (written with BGYM_FLAG_RANDOM_POLICY, nullable)
  const BgymDraws* draws;    // n (nullable)
  uint8_t* obs;              // n x 176 (nullable)
  uint8_t* sel;              // n x 16  selection records (BgymSel), nullable together with obs
  uint8_t* obs_dirty;        // n (nullable): |= BGYM_OBS_DIRTY* for every env whose observation record is rewritten
  double* reward;            // n
  uint8_t* terminated;       // n
  uint8_t* truncated;        // n (nullable)
  BgymInfo* info;            // n (nullable)
  // reset-only
  const uint8_t* reset_mask; // n (nullable)
  const uint32_t* seeds;     // n
  const uint8_t* decks52;    // n x 52 (nullable)
  long long n;
  int flags;
  // partitioned step (bgym_step_part.cuh): device lists of deferred env indices + their counters
  int* part_lists;           // [N_LISTS][part_cap]
  int* part_counters;        // [N_LISTS * PART_CTR_STRIDE] (counter of list l at l * PART_CTR_STRIDE)
  uint16_t* part_aux;        // [part_cap] per env: draw-sequence position of a step continued at level 2
  long long part_cap;
};

}  // namespace bgym
#ifndef BGYM_SMALL_N_DEFAULT
#define BGYM_SMALL_N_DEFAULT 65536
#endif
#include "bgym_step_part.cuh"
#include "bgym_rollout.cuh"
namespace bgym {

// ---------------------------------------------------------------------------------------------
// K3: reset.  Per-warp tiles; with a reset mask the untouched envs are loaded and re-emitted.
// ---------------------------------------------------------------------------------------------
constexpr int RESET_WARPS = 4;
constexpr int RESET_WARP_SMEM = 32 * (BGYM_HOT_BYTES + BGYM_COLD_BYTES + BGYM_OBS_BYTES);
constexpr int RESET_CTA_SMEM = RESET_WARPS * RESET_WARP_SMEM + 16 * RESET_WARPS;

__global__ void __launch_bounds__(RESET_WARPS * 32, 3) env_reset_kernel(StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* hot_buf = smem + warp * RESET_WARP_SMEM;
  uint8_t* cold_buf = hot_buf + 32 * BGYM_HOT_BYTES;
  uint8_t* obs_buf = cold_buf + 32 * BGYM_COLD_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + RESET_WARPS * RESET_WARP_SMEM) + warp * 2;
  if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncwarp();
  const long long n_tiles = (a.n + 31) >> 5;
  const long long warp_gid = (long long)blockIdx.x * RESET_WARPS + warp;
  const long long warp_cnt = (long long)gridDim.x * RESET_WARPS;
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  const bool need_load = a.reset_mask != nullptr;   // a reset of ALL envs needs no state load
  uint32_t parity = 0;
  for (long long tile = warp_gid; tile < n_tiles; tile += warp_cnt) {
    const long long e = tile * 32 + lane;
    const bool active = e < a.n;
    const uint32_t cnt = (uint32_t)min(32LL, a.n - tile * 32);
    uint8_t* hot = hot_buf + lane * BGYM_HOT_BYTES;
    uint8_t* cold = cold_buf + lane * BGYM_COLD_BYTES;
    uint8_t* obs_s = obs_buf + lane * BGYM_OBS_BYTES;
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (need_load) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, cnt * (BGYM_HOT_BYTES + BGYM_COLD_BYTES));
        bulk_g2s(hot_buf, a.hot + tile * 32 * BGYM_HOT_BYTES, cnt * BGYM_HOT_BYTES, bar);
        bulk_g2s(cold_buf, a.cold + tile * 32 * BGYM_COLD_BYTES, cnt * BGYM_COLD_BYTES, bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
    }
    if (active) {
      Hot h;
      if (a.reset_mask && !a.reset_mask[e]) {
        unpack_hot(hot, h);   // untouched env: only re-emit its observation; its toggle record is the current one
        set_chunk1(h, __ldcg(reinterpret_cast<const uint4*>(a.tog + e * BGYM_TOG_BYTES)));
        pack_hot(hot, h);
      } else {
        reset_hot(h, a.seeds[e]);
        if (a.flags & BGYM_FLAG_GEN_C3) gen_hot(h, a.seeds[e], a.flags);
        reset_blocks_serial(cold, a.seeds[e], a.decks52 ? a.decks52 + e * 52 : nullptr, (a.flags & BGYM_FLAG_GEN_C3) != 0);
        pack_hot(hot, h);
        hot_clear_extra(hot);
      }
      store_tog(a.tog + e * BGYM_TOG_BYTES, h);
      if (with_obs) {
        const uint64_t m = action_mask(h, cold);
        write_obs(h, cold, m, obs_s);
        *reinterpret_cast<uint4*>(a.sel + e * BGYM_SEL_BYTES) = sel_words(h, m);
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(a.hot + tile * 32 * BGYM_HOT_BYTES, hot_buf, cnt * BGYM_HOT_BYTES);
      bulk_s2g(a.cold + tile * 32 * BGYM_COLD_BYTES, cold_buf, cnt * BGYM_COLD_BYTES);
      if (with_obs) bulk_s2g(a.obs + tile * 32 * BGYM_OBS_BYTES, obs_buf, cnt * BGYM_OBS_BYTES);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait0();
}

// ---------------------------------------------------------------------------------------------
// K1: hand scoring.  One thread per hand; inputs are dense per-hand arrays.
// ---------------------------------------------------------------------------------------------
struct ScoreArgs {
  const uint8_t* cards8; const uint16_t* mods8; const uint8_t* n_cards; const uint8_t* jokers8;
  const uint8_t* levels12; const BgymScoreCtx* ctx;
  uint8_t* hand_type; int32_t* chips; int32_t* mult; double* x_mult; long long* score; int32_t* money;
  uint32_t seed; long long n; int flags;
};

// does the context hand name equal the table's name? (complete_joker_effects.py:64-80 vs
// balatro_env_2.py:674 — 'Pair' / 'Three of a Kind' / 'Four of a Kind' only with table naming)
__device__ __forceinline__ bool name_matches(int ht, bool table_names, int hn) {
  switch (hn) {
    case BGYM_HN_PAIR: return table_names && ht == BGYM_HT_ONE_PAIR;
    case BGYM_HN_THREE_OAK: return table_names && ht == BGYM_HT_THREE_KIND;
    case BGYM_HN_FOUR_OAK: return table_names && ht == BGYM_HT_FOUR_KIND;
    case BGYM_HN_TWO_PAIR: return ht == BGYM_HT_TWO_PAIR;
    case BGYM_HN_STRAIGHT: return ht == BGYM_HT_STRAIGHT;
    case BGYM_HN_FLUSH: return ht == BGYM_HT_FLUSH;
  }
  return false;
}

struct ScoreRng {  // native draws of the scoring path: Philox keyed by (seed), counter (blk, hand index, 1)
  uint32_t seed, ctr; unsigned long long index; uint4 buf; int pos, skip;   // skip: words of the next block already consumed
  __device__ __forceinline__ uint32_t word() {
    if (pos == 4) { buf = philox4x32_10(ctr++, (uint32_t)index, (uint32_t)(index >> 32), 1, seed, BGYM_PHILOX_KEY1); pos = skip; skip = 0; }
    uint32_t w = pos == 0 ? buf.x : pos == 1 ? buf.y : pos == 2 ? buf.z : buf.w;
    pos++;
    return w;
  }
  __device__ __forceinline__ double u01() {
    uint32_t x = word() >> 5; uint32_t y = word() >> 6;
    return (x * 67108864.0 + y) * (1.0 / 9007199254740992.0);
  }
  __device__ __forceinline__ int below(int n) {
    uint32_t un = (uint32_t)n;
    uint64_t m = (uint64_t)word() * un; uint32_t l = (uint32_t)m;
    if (l < un) { uint32_t t = (0u - un) % un; while (l < t) { m = (uint64_t)word() * un; l = (uint32_t)m; } }
    return (int)(m >> 32);
  }
};

// Fast path of config 2 (five-card plays, no modifiers, no jokers, level-1 hands).  The kernel is
// bound by instruction issue before HBM (one thread per hand, 32 B per hand), so classification is
// done on 13-bit rank masks instead of the general 4-bit-per-rank histogram: s_k = ranks seen at
// least k times, updated with four LOP3 per card; flush by XOR-ing the packed codes against card 0;
// a straight is five distinct ranks whose mask is 31 << lowest, or the wheel.  Same priority order
// as classify() (balatro_game.py:40-93) for ANY five codes, duplicates included.
__device__ __forceinline__ void score_hand5(const ScoreArgs& a, long long i, const uint2 cw) {
  const uint32_t c0 = cw.x & 0xFF, c1 = (cw.x >> 8) & 0xFF, c2 = (cw.x >> 16) & 0xFF, c3 = cw.x >> 24, c4 = cw.y & 0xFF;
  const uint32_t m1 = 1u << (c1 >> 2), m2 = 1u << (c2 >> 2), m3 = 1u << (c3 >> 2), m4 = 1u << (c4 >> 2);
  uint32_t s1 = 1u << (c0 >> 2), s2 = 0, s3 = 0, s4 = 0;
  s2 |= s1 & m1; s1 |= m1;
  s3 |= s2 & m2; s2 |= s1 & m2; s1 |= m2;
  s4 |= s3 & m3; s3 |= s2 & m3; s2 |= s1 & m3; s1 |= m3;
  s4 |= s3 & m4; s3 |= s2 & m4; s2 |= s1 & m4; s1 |= m4;
  const bool flush = ((((cw.x ^ (c0 * 0x01010101u)) & 0x03030303u) | ((cw.y ^ c0) & 3u)) == 0u);
  const uint32_t low = s1 & (0u - s1);
  const bool straight = (s2 == 0u) && (s1 == low * 31u || s1 == 0x100Fu);
  const int n2 = __popc(s2);
  int ht = n2 == 1 ? BGYM_HT_ONE_PAIR : BGYM_HT_HIGH_CARD;
  ht = n2 == 2 ? BGYM_HT_TWO_PAIR : ht;
  ht = s3 ? BGYM_HT_THREE_KIND : ht;
  ht = straight ? BGYM_HT_STRAIGHT : ht;
  ht = flush ? BGYM_HT_FLUSH : ht;
  ht = (s3 && n2 == 2) ? BGYM_HT_FULL_HOUSE : ht;
  ht = s4 ? BGYM_HT_FOUR_KIND : ht;
  ht = (straight && flush) ? BGYM_HT_STRAIGHT_FLUSH : ht;
  const int chip_sum = card_chips(c0, 0, 0) + card_chips(c1, 0, 0) + card_chips(c2, 0, 0) + card_chips(c3, 0, 0) + card_chips(c4, 0, 0);
  const int chips = c_base_chips[ht] + chip_sum, mult = c_base_mult[ht];
  a.hand_type[i] = (uint8_t)ht;
  a.chips[i] = chips;
  a.mult[i] = mult;
  a.score[i] = (long long)chips * mult;   // x_mult == 1.0: int(chips * mult * 1.0)
  if (a.x_mult) a.x_mult[i] = 1.0;
  if (a.money) a.money[i] = 0;
}

// four hands per thread and iteration: the four 8-byte loads are issued before any of them is used,
// so a resident warp keeps 1 KB in flight instead of 256 B (the pass was load-latency bound)
constexpr int HANDS5_UNROLL = 4;
__global__ void __launch_bounds__(256) score_hands5_kernel(ScoreArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const uint2* src = reinterpret_cast<const uint2*>(a.cards8);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (HANDS5_UNROLL - 1) * stride < a.n; i += HANDS5_UNROLL * stride) {
    uint2 cw[HANDS5_UNROLL];
#pragma unroll
    for (int u = 0; u < HANDS5_UNROLL; u++) cw[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < HANDS5_UNROLL; u++) score_hand5(a, i + u * stride, cw[u]);
  }
  for (; i < a.n; i += stride) score_hand5(a, i, __ldg(src + i));
}

// One joker-effect row as registers (BgymJokerFx is 16 bytes: kind u8 | pad | arg u16 || chips i16 | mult i16 || xmult f32 || money i16 | pad)
struct FxRow {
  uint4 q;
  __device__ __forceinline__ int kind() const { return (int)(q.x & 0xFF); }
  __device__ __forceinline__ int arg() const { return (int)(q.x >> 16); }
  __device__ __forceinline__ int chips() const { return (int)(short)(q.y & 0xFFFF); }
  __device__ __forceinline__ int mult() const { return (int)(short)(q.y >> 16); }
  __device__ __forceinline__ float xmult() const { return __uint_as_float(q.z); }
  __device__ __forceinline__ int money() const { return (int)(short)(q.w & 0xFFFF); }
};

// The interpreter's tables are staged in SHARED memory per CTA: per-lane joker ids index them divergently, which the
// constant cache serialises (one replay per distinct address in the warp) — in the (card x joker) loop that was ~99 %
// of the first version's time.  Three words per joker id:
//   s_fx   the 16-byte effect row (chips, mult, xmult, money)
//   s_jm   individual phase: the joker as a mask over card bits {rank 0..14} u {16 + suit 0..4}, bit 31 = Bloodstone
//   s_uop  main phase as ONE branch-free micro-op: which hand predicate gates the joker, which of its three effects
//          apply, and what scales them — every lane of a warp executes the same ~25 instructions per joker slot whatever
//          its joker is (the switch over effect kinds ran once per distinct kind in the warp: 14 % of the kernel's
//          instructions at 1-3 active lanes, ncu)
// hand predicates (bit index in the per-hand word P)
enum { HP_ALWAYS = 0, HP_SUIT0 = 1 /* ..4 */, HP_HALF = 5, HP_LAST_HAND = 6, HP_NO_DISCARDS = 7, HP_ALL_BLACK = 8, HP_SEEING_DOUBLE = 9,
       HP_FLOWER_POT = 10, HP_KINGS = 11, HP_QUEENS = 12, HP_NAME0 = 13 /* + BGYM_HN_* (6 names) */, HP_NEVER = 31 };
// micro-op fields
enum { UOP_PRED_BITS = 5, UOP_CSCALE_SHIFT = 5, UOP_MSCALE_SHIFT = 7, UOP_CHIPS = 1 << 9, UOP_MULT = 1 << 10, UOP_XMULT = 1 << 11,
       UOP_BARON = 1 << 12, UOP_MISPRINT = 1 << 13 };
enum { SC_ONE = 0, SC_DISCARDS = 1, SC_DECK = 2, SM_ONE = 0, SM_NJ = 1, SM_QUEENS = 2 };
__device__ __forceinline__ uint32_t main_uop(int kind, int arg) {
  switch (kind) {
    case BGYM_FX_MAIN_ALWAYS: return HP_ALWAYS | UOP_CHIPS | UOP_MULT | UOP_XMULT;
    case BGYM_FX_MAIN_HANDNAME: return (arg >= 0 && arg < 6 ? HP_NAME0 + arg : HP_NEVER) | UOP_CHIPS | UOP_MULT | UOP_XMULT;
    case BGYM_FX_MAIN_SUIT_ANY: return (HP_SUIT0 + (arg & 3)) | UOP_MULT;
    case BGYM_FX_MAIN_STATE:
      switch (arg) {
        case BGYM_FXS_HALF: return HP_HALF | UOP_MULT;
        case BGYM_FXS_ABSTRACT: return HP_ALWAYS | UOP_MULT | (SM_NJ << UOP_MSCALE_SHIFT);
        case BGYM_FXS_ACROBAT: return HP_LAST_HAND | UOP_XMULT;
        case BGYM_FXS_MYSTIC: return HP_NO_DISCARDS | UOP_MULT;
        case BGYM_FXS_BANNER: return HP_ALWAYS | UOP_CHIPS | (SC_DISCARDS << UOP_CSCALE_SHIFT);
        case BGYM_FXS_BLUE: return HP_ALWAYS | UOP_CHIPS | (SC_DECK << UOP_CSCALE_SHIFT);
        case BGYM_FXS_MISPRINT: return HP_NEVER | UOP_MISPRINT;
      }
      return HP_NEVER;
    case BGYM_FX_MAIN_SPECIAL:
      switch (arg) {
        case BGYM_FXSP_BLACKBOARD: return HP_ALL_BLACK | UOP_XMULT;
        case BGYM_FXSP_SEEING_DOUBLE: return HP_SEEING_DOUBLE | UOP_XMULT;
        case BGYM_FXSP_FLOWER_POT: return HP_FLOWER_POT | UOP_XMULT;
        case BGYM_FXSP_BARON: return HP_KINGS | UOP_XMULT | UOP_BARON;
        case BGYM_FXSP_SHOOT_MOON: return HP_QUEENS | UOP_MULT | (SM_QUEENS << UOP_MSCALE_SHIFT);
      }
      return HP_NEVER;
  }
  return HP_NEVER;    // individual-phase jokers and empty rows do nothing in the main phase
}
__device__ __forceinline__ uint32_t ind_mask(int kind, int arg) {
  return (kind == BGYM_FX_IND_RANKSET || kind == BGYM_FX_IND_FACE) ? (uint32_t)arg
       : (kind == BGYM_FX_IND_SUIT) ? ((1u << (16 + (arg & 3))) | ((arg & 0x80) ? 0x80000000u : 0u)) : 0u;
}

__global__ void __launch_bounds__(256) score_hands_kernel(ScoreArgs a) {
  __shared__ uint4 s_fx[BGYM_NUM_JOKERS + 1];
  __shared__ uint32_t s_jm[BGYM_NUM_JOKERS + 1], s_uop[BGYM_NUM_JOKERS + 1];
  for (int t = threadIdx.x; t < BGYM_NUM_JOKERS + 1; t += blockDim.x) {
    const BgymJokerFx f = c_joker_fx[t];
    uint4 q;
    q.x = (uint32_t)f.kind | ((uint32_t)f.arg << 16);
    q.y = (uint32_t)(uint16_t)f.chips | ((uint32_t)(uint16_t)f.mult << 16);
    q.z = __float_as_uint(f.xmult);
    q.w = (uint32_t)(uint16_t)f.money;
    s_fx[t] = q;
    s_jm[t] = t ? ind_mask(f.kind, f.arg) : 0u;
    s_uop[t] = t ? main_uop(f.kind, f.arg) : (uint32_t)HP_NEVER;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const bool table_names = (a.flags & BGYM_SCORE_TABLE_NAMES) != 0;
  // warp-uniform trip count: the Bloodstone rolls below are made by the whole warp
  for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < a.n; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + lane;
    const bool valid = i < a.n;
    uint2 cw = make_uint2(0, 0);
    if (valid) cw = __ldg(reinterpret_cast<const uint2*>(a.cards8) + i);
    uint64_t cards = u64_of(cw.x, cw.y);
    int nc = !valid ? 0 : (a.n_cards ? a.n_cards[i] : 5);
    uint4 mw = make_uint4(0, 0, 0, 0);
    if (a.mods8 && valid) mw = __ldg(reinterpret_cast<const uint4*>(a.mods8) + i);
    // classification + card chips
    HandHist hist;
    hist.clear();
    int chip_sum = 0;
    uint32_t suit_present = 0, stone_present = 0, all_black = 1;
    int kings = 0, queens = 0;
    // per-card (rank, suit) as seen by the joker tables: stone cards show rank 0 / suit 'Stone'
    uint64_t rank8 = 0; uint32_t suit8 = 0;  // suit nibble 4 = Stone
    for (int c = 0; c < nc; c++) {
      int code = byte_at(cards, c);
      uint32_t mword = c < 2 ? mw.x : c < 4 ? mw.y : c < 6 ? mw.z : mw.w;
      int m = (mword >> (16 * (c & 1))) & 0xFFFF;
      int enh = m & 15, ed = (m >> 4) & 15;
      hist.add(code);
      chip_sum += card_chips(code, enh, ed);
      bool stone = enh == BGYM_ENH_STONE;
      int rank = stone ? 0 : (code >> 2) + 2;
      int suit = stone ? 4 : (code & 3);
      rank8 |= (uint64_t)rank << (8 * c);
      suit8 |= (uint32_t)suit << (4 * c);
      if (stone) stone_present = 1; else suit_present |= 1u << suit;
      if (!(suit == 3 || suit == 0)) all_black = 0;
      kings += rank == 13; queens += rank == 12;
    }
    const bool with_jokers = a.jokers8 != nullptr;
    uint64_t jraw = 0;
    if (with_jokers && valid) {
      const uint2 jw = __ldg(reinterpret_cast<const uint2*>(a.jokers8) + i);
      jraw = u64_of(jw.x, jw.y);
    }
    int ht;
    if (a.flags & BGYM_SCORE_RULES) {      // balatro_sim.py:220-400; Four Fingers / Shortcut are read from the joker slots
      uint32_t scnt = 0;
      for (int c = 0; c < nc; c++) scnt += 1u << (8 * (byte_at(cards, c) & 3));
      bool four_fingers = false, shortcut = false;
      for (int j = 0; j < 8; j++) { const int id = byte_at(jraw, j); four_fingers |= id == BGYM_J_FOUR_FINGERS; shortcut |= id == BGYM_J_SHORTCUT; }
      ht = classify_rules(hist.cnt, scnt, hist.rmask, nc, four_fingers, shortcut);
    } else {
      ht = classify(hist);
    }
    int lvl = (a.levels12 && valid) ? a.levels12[i * 12 + ht] : 1;
    int chips, mult;
    hand_base(ht, lvl, chips, mult);
    chips += chip_sum;
    double x_mult = 1.0;
    int money = 0;
    if (with_jokers) {      // (kernel argument: the whole warp takes this branch)
      // K4: table-driven joker interpreter (unified_scoring.py:156-244), joker order preserved
      uint64_t jk = 0;
      int nj = 0;
      for (int j = 0; j < 8; j++) { int id = byte_at(jraw, j); if (id) { jk |= (uint64_t)min(id, BGYM_NUM_JOKERS) << (8 * nj); nj++; } }
      BgymScoreCtx cx;
      cx.hands_left = 4; cx.discards_left = 3; cx.deck_len = 52; cx.use_replay = 0; cx.bloodstone_bits = 0;
      cx.misprint[0] = 0;
      if (a.ctx && valid) cx = a.ctx[i];
      ScoreRng rng; rng.seed = a.seed; rng.ctr = 0; rng.index = (unsigned long long)i; rng.pos = 4; rng.skip = 0;
      // individual phase: card-major, joker-minor (:173-209).  A card is a bit set {rank 0..14} u {16 + suit 0..4};
      // a rank-set / face / suit joker is a mask over the same bits (s_jm), so "fires" is one AND.  Bit 31 marks
      // Bloodstone, which rolls for every card whether the suit matches or not.
      uint32_t jm[8];
      int n_blood = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        jm[j] = j < nj ? s_jm[byte_at(jk, j)] : 0u;
        n_blood += jm[j] >> 31;
      }
      // Bloodstone rolls (:161, `random() < 0.5` per (card, joker) pair in loop order): roll k reads stream words 2k and
      // 2k + 1, and u01 < 0.5 is "bit 31 of word 2k is clear" (u01 = ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53).  Philox blocks
      // are independent, so the WARP makes them: lane b computes block b of the hand that needs rolls — one converged
      // Philox pass per such hand instead of a chain of up to 32 on one lane (16 % of the kernel's instructions, ncu).
      const int n_rolls = cx.use_replay ? 0 : nc * n_blood;
      uint32_t roll_even = 0, roll_odd = 0;      // bit b: roll 2b / roll 2b + 1 hit
      {
        uint32_t need = __ballot_sync(0xffffffffu, n_rolls > 0);
        while (need) {
          const int src = __ffs((int)need) - 1;
          need &= need - 1;
          const unsigned long long idx = __shfl_sync(0xffffffffu, rng.index, src);
          const uint4 blk = philox4x32_10_inl((uint32_t)lane, (uint32_t)idx, (uint32_t)(idx >> 32), 1, a.seed, BGYM_PHILOX_KEY1);
          const uint32_t e = __ballot_sync(0xffffffffu, !(blk.x >> 31)), o = __ballot_sync(0xffffffffu, !(blk.z >> 31));
          if (lane == src) { roll_even = e; roll_odd = o; }
        }
      }
      int ind_chips = 0, ind_mult = 0, roll = 0; double ind_x = 1.0;
#pragma unroll 1
      for (int c = 0; c < nc; c++) {
        const uint32_t cardbits = (1u << byte_at(rank8, c)) | (1u << (16 + nib_at(suit8, c))) | 0x80000000u;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (cardbits & jm[j]) {          // rare: the card matches, or the joker is Bloodstone
            bool fire = (cardbits & jm[j] & 0x7FFFFFFFu) != 0u;
            if (jm[j] & 0x80000000u) {
              const bool hit = cx.use_replay ? ((cx.bloodstone_bits >> c) & 1) : ((((roll & 1) ? roll_odd : roll_even) >> (roll >> 1)) & 1u);
              roll++;
              fire = fire && hit;
            }
            if (fire) {
              const FxRow fx = {s_fx[byte_at(jk, j)]};
              ind_chips += fx.chips(); ind_mult += fx.mult(); money += fx.money(); ind_x *= (double)fx.xmult();
            }
          }
        }
      }
      rng.ctr = (uint32_t)(n_rolls >> 1); rng.skip = (n_rolls & 1) * 2;     // the stream stands behind word 2 * n_rolls
      chips += ind_chips; mult += ind_mult; x_mult *= ind_x;
      // main phase, joker order (:211-244): one micro-op per joker slot (s_uop)
      const int n_suit_names = __popc(suit_present) + (int)stone_present;
      uint32_t P = 1u | (suit_present & 15u) << HP_SUIT0;
      P |= (nc <= 3 ? 1u : 0u) << HP_HALF | (cx.hands_left == 1 ? 1u : 0u) << HP_LAST_HAND | (cx.discards_left == 0 ? 1u : 0u) << HP_NO_DISCARDS;
      P |= (all_black ? 1u : 0u) << HP_ALL_BLACK | (((suit_present & 1) && n_suit_names > 1) ? 1u : 0u) << HP_SEEING_DOUBLE;
      P |= (n_suit_names == 4 ? 1u : 0u) << HP_FLOWER_POT | (kings > 0 ? 1u : 0u) << HP_KINGS | (queens > 0 ? 1u : 0u) << HP_QUEENS;
      {  // hand-name jokers (complete_joker_effects.py:64-80 vs balatro_env_2.py:674)
        const uint32_t tn = table_names ? 1u : 0u;
        P |= (tn & (ht == BGYM_HT_ONE_PAIR)) << (HP_NAME0 + BGYM_HN_PAIR) | (tn & (ht == BGYM_HT_THREE_KIND)) << (HP_NAME0 + BGYM_HN_THREE_OAK);
        P |= (tn & (ht == BGYM_HT_FOUR_KIND)) << (HP_NAME0 + BGYM_HN_FOUR_OAK) | (uint32_t)(ht == BGYM_HT_TWO_PAIR) << (HP_NAME0 + BGYM_HN_TWO_PAIR);
        P |= (uint32_t)(ht == BGYM_HT_STRAIGHT) << (HP_NAME0 + BGYM_HN_STRAIGHT) | (uint32_t)(ht == BGYM_HT_FLUSH) << (HP_NAME0 + BGYM_HN_FLUSH);
      }
      const uint64_t cscale = 1ull | (uint64_t)cx.discards_left << 16 | (uint64_t)cx.deck_len << 32;
      const uint64_t mscale = 1ull | (uint64_t)nj << 16 | (uint64_t)queens << 32;
      const double baron = c_pow_1_5[min(kings, 100)];
      int misprint_seen = 0;
#pragma unroll 1
      for (int j = 0; j < nj; j++) {
        const int id = byte_at(jk, j);
        const uint32_t u = s_uop[id];
        const FxRow row = {s_fx[id]};
        const bool on = (P >> (u & ((1u << UOP_PRED_BITS) - 1))) & 1u;
        const int cs = (int)((cscale >> (16 * ((u >> UOP_CSCALE_SHIFT) & 3u))) & 0xFFFFu);
        const int ms = (int)((mscale >> (16 * ((u >> UOP_MSCALE_SHIFT) & 3u))) & 0xFFFFu);
        int ec = (on && (u & UOP_CHIPS)) ? row.chips() * cs : 0;
        int em = (on && (u & UOP_MULT)) ? row.mult() * ms : 0;
        double ex = (on && (u & UOP_XMULT)) ? ((u & UOP_BARON) ? baron : (double)row.xmult()) : 1.0;
        if (u & UOP_MISPRINT) {      // the one main-phase effect that draws (rare)
          em = cx.use_replay ? cx.misprint[min(misprint_seen, 4)] : rng.below(24);
          misprint_seen++;
        }
        chips += ec; mult += em; x_mult *= ex;
      }
    }
    // final_score = int(chips * mult * x_mult), unified_scoring.py:286
    long long score = (long long)((double)((long long)chips * (long long)mult) * x_mult);
    if (valid) {
      a.hand_type[i] = (uint8_t)ht;
      a.chips[i] = chips;
      a.mult[i] = mult;
      if (a.x_mult) a.x_mult[i] = x_mult;
      a.score[i] = score;
      if (a.money) a.money[i] = money;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void action_mask_kernel(const uint8_t* hot, const uint8_t* tog, const uint8_t* cold, uint64_t* mask, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* hr = hot + i * BGYM_HOT_BYTES;
    const uint8_t* tr = tog + i * BGYM_TOG_BYTES;    // hand_n / sel_n / phase live in the toggle record (BgymTog)
    const uint8_t* cr = cold + i * BGYM_COLD_BYTES;
    uint64_t m = 0;
    int phase = tr[9];
    if (phase == BGYM_PHASE_PLAY) {
      int hand_n = tr[0], sel_n = tr[2], discards_left = hr[129], cons_n = hr[131];
      m = ((1ull << min(hand_n, 8)) - 1) << BGYM_A_SELECT_BASE;
      if (sel_n > 0) m |= 1ull << BGYM_A_PLAY_HAND;
      if (sel_n > 0 && discards_left > 0) m |= 1ull << BGYM_A_DISCARD;
      m |= ((1ull << cons_n) - 1) << BGYM_A_USE_CONS_BASE;
    } else if (phase == BGYM_PHASE_SHOP) {
      int money = *reinterpret_cast<const int*>(hr + 40), joker_n = hr[130];
      int n_items = cr[OFF_N_ITEMS];
      for (int k = 0; k < n_items; k++)
        if (money >= *reinterpret_cast<const int*>(cr + OFF_ITEM_COST + 4 * k)) m |= 1ull << (BGYM_A_SHOP_BUY_BASE + k);
      if (money >= *reinterpret_cast<const int*>(hr + 116)) m |= 1ull << BGYM_A_SHOP_REROLL;
      m |= 1ull << BGYM_A_SHOP_END;
      m |= ((1ull << joker_n) - 1) << BGYM_A_SELL_JOKER_BASE;
    } else if (phase == BGYM_PHASE_BLIND_SELECT) {
      m = 0xFull << BGYM_A_SELECT_BLIND_BASE;
    }
    mask[i] = m;
  }
}


// uniform random legal action from obs.action_mask_bits; Philox keyed (seed, policy key),
// counter (env index, step)
// step_dev != nullptr: the step number lives on the device (launches replayed from a CUDA graph cannot take it as
// a kernel argument); bump_counter_kernel advances it after the sampler
__global__ void bump_counter_kernel(unsigned long long* ctr) { *ctr += 1; }
__global__ void sample_actions_kernel(const uint8_t* mask_words, long long mask_stride, int32_t* actions, uint32_t seed, unsigned long long step,
                                      const unsigned long long* step_dev, long long n) {
  if (step_dev) step += *step_dev;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint64_t m = *reinterpret_cast<const uint64_t*>(mask_words + i * mask_stride);
    int cnt = __popcll(m);
    int act = 0;
    if (cnt) {
      uint4 w = philox4x32_10((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), (uint32_t)step, (uint32_t)(step >> 32), seed, BGYM_POLICY_KEY1);
      int k = (int)__umulhi(w.x, (uint32_t)cnt);
      act = select_bit64(m, k);
    }
    actions[i] = act;
  }
}

// K6: fold one step's (reward, terminated) into per-env accumulators and slab statistics
// stats[0]=episodes [1]=sum return [2]=sum length [3]=steps [4]=sum reward
__global__ void episode_stats_kernel(const double* reward, const uint8_t* terminated, double* ret_acc, uint32_t* len_acc,
                                     double* stats, long long n) {
  double eps = 0, sret = 0, slen = 0, steps = 0, srew = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double r = reward[i];
    double acc = ret_acc[i] + r;
    uint32_t len = len_acc[i] + 1;
    steps += 1; srew += r;
    if (terminated[i]) { eps += 1; sret += acc; slen += len; acc = 0; len = 0; }
    ret_acc[i] = acc; len_acc[i] = len;
  }
  // warp reduce, then block reduce through shared memory: one atomic per CTA and statistic
  for (int o = 16; o > 0; o >>= 1) {
    eps += __shfl_xor_sync(0xffffffffu, eps, o); sret += __shfl_xor_sync(0xffffffffu, sret, o);
    slen += __shfl_xor_sync(0xffffffffu, slen, o); steps += __shfl_xor_sync(0xffffffffu, steps, o);
    srew += __shfl_xor_sync(0xffffffffu, srew, o);
  }
  __shared__ double part[8][5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { part[warp][0] = eps; part[warp][1] = sret; part[warp][2] = slen; part[warp][3] = steps; part[warp][4] = srew; }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += part[w][threadIdx.x];
    if (v != 0) atomicAdd(&stats[threadIdx.x], v);
  }
}

// ---------------------------------------------------------------------------------------------
// coherence between the record arrays and their side arrays (BgymTog / BgymSel, include/bgym.h)
// ---------------------------------------------------------------------------------------------
__global__ void sync_state_kernel(uint8_t* hot, uint8_t* tog, long long n, int from_records) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint8_t* hr = hot + i * BGYM_HOT_BYTES;
    uint8_t* tr = tog + i * BGYM_TOG_BYTES;
    if (from_records) {
      Hot h;
      unpack_hot(hr, h);
      store_tog(tr, h);
    } else {
      *reinterpret_cast<uint4*>(hr + 16) = *reinterpret_cast<const uint4*>(tr);
    }
  }
}
__global__ void sync_obs_kernel(uint8_t* obs, uint8_t* sel, long long n, int from_records) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint8_t* o = obs + i * BGYM_OBS_BYTES;
    uint8_t* s = sel + i * BGYM_SEL_BYTES;
    if (from_records) {
      const uint2 sc = *reinterpret_cast<const uint2*>(o + 8), mk = *reinterpret_cast<const uint2*>(o + 160);
      *reinterpret_cast<uint4*>(s) = make_uint4(sc.x, sc.y, mk.x, mk.y);
    } else {
      const uint4 q = *reinterpret_cast<const uint4*>(s);
      *reinterpret_cast<uint2*>(o + 8) = make_uint2(q.x, q.y);
      *reinterpret_cast<uint2*>(o + 160) = make_uint2(q.z, q.w);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// observation deltas for a host mirror: the records a step rewrote are exactly those of the envs on its level-1
// lists.  pack: gather them (index + record) into a dense staging block; scatter: write staged records to their
// places in another array of BgymObs — pinned host memory included (zero-copy stores over PCIe).
// staging: int32 count @0 | int32 index[cap] @16 | BgymObs[cap] @(16 + 4 cap, rounded up to 16)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t dirty_rec_offset(long long cap) { return 16 + (((size_t)cap * 4 + 15) & ~(size_t)15); }
// A staged record is bytes 0..159 of the observation record (BGYM_OBS_DELTA_BYTES): bytes 160..175 are the mask word,
// which travels in the selection record, and padding.  Bit 31 of a staged index says that the record's SHOP CHUNKS
// (6, 7: shop_items[1..9], shop_costs[0..6]) may differ from what the mirror holds (BGYM_OBS_DIRTY_SHOP).
// The records are staged IN ASCENDING ENV ORDER: a stream compaction of the per-env dirty bytes the step wrote
// (count per block of PACK_ENVS envs -> exclusive prefix -> indices), then a coalesced record move, 10 lanes per record.
// Ascending order matters to the zero-copy scatter that follows: 128-byte stores into pinned host memory ran at
// 51.5 GB/s in address order and 42.4 GB/s in random order (tools/exp/zc_probe.cu).
constexpr int DELTA_LANES = BGYM_OBS_DELTA_BYTES / 16, DELTA_PER_WARP = 32 / DELTA_LANES;
constexpr int PACK_THREADS = 256, PACK_PER_THREAD = 16, PACK_ENVS = PACK_THREADS * PACK_PER_THREAD;
__device__ __forceinline__ int dirty_count16(const uint4 f) {      // number of bytes with BGYM_OBS_DIRTY set
  return __popc(f.x & 0x01010101u) + __popc(f.y & 0x01010101u) + __popc(f.z & 0x01010101u) + __popc(f.w & 0x01010101u);
}
__device__ __forceinline__ uint4 load_flags16(const uint8_t* dirty, long long e0, long long n, bool all) {
  if (all) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      w[q] = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) if (e0 + 4 * q + k < n) w[q] |= (uint32_t)(BGYM_OBS_DIRTY | BGYM_OBS_DIRTY_SHOP) << (8 * k);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (e0 + 16 <= n) return *reinterpret_cast<const uint4*>(dirty + e0);
  uint32_t w[4] = {0, 0, 0, 0};
  for (int k = 0; k < 16 && e0 + k < n; k++) w[k >> 2] |= (uint32_t)dirty[e0 + k] << (8 * (k & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void __launch_bounds__(PACK_THREADS) pack_count_kernel(const uint8_t* __restrict__ dirty, int* __restrict__ counts, long long n, int all) {
  const long long e0 = ((long long)blockIdx.x * PACK_THREADS + threadIdx.x) * PACK_PER_THREAD;
  int c = e0 < n ? dirty_count16(load_flags16(dirty, e0, n, all != 0)) : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ int part[PACK_THREADS / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < PACK_THREADS / 32; w++) t += part[w];
    counts[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(PACK_THREADS) pack_index_kernel(uint8_t* __restrict__ dirty, const int* __restrict__ counts,
                                                                  uint8_t* __restrict__ staging, long long cap, long long n, int all) {
  __shared__ int part[PACK_THREADS / 32];
  __shared__ int s_base;
  // exclusive prefix of the block counts before this block
  int b = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += PACK_THREADS) b += counts[i];
  for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < PACK_THREADS / 32; w++) t += part[w];
    s_base = t;
    if (blockIdx.x == gridDim.x - 1) *reinterpret_cast<int*>(staging) = t + counts[blockIdx.x];
  }
  __syncthreads();
  const long long e0 = ((long long)blockIdx.x * PACK_THREADS + threadIdx.x) * PACK_PER_THREAD;
  const uint4 f = e0 < n ? load_flags16(dirty, e0, n, all != 0) : make_uint4(0, 0, 0, 0);
  const int c = dirty_count16(f);
  // block-exclusive prefix of c
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += v; }
  __syncthreads();
  if ((threadIdx.x & 31) == 31) part[threadIdx.x >> 5] = incl;
  __syncthreads();
  int off = s_base + incl - c;
  for (int w = 0; w < (int)(threadIdx.x >> 5); w++) off += part[w];
  uint32_t* idx_out = reinterpret_cast<uint32_t*>(staging + 16);
  const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const uint32_t fl = (fw[k >> 2] >> (8 * (k & 3))) & 0xFF;
    if (fl & BGYM_OBS_DIRTY) {
      if (off < cap) idx_out[off] = (uint32_t)(e0 + k) | ((fl & BGYM_OBS_DIRTY_SHOP) ? 0x80000000u : 0u);
      off++;
    }
  }
  if (!all && c && e0 + 16 <= n) *reinterpret_cast<uint4*>(dirty + e0) = make_uint4(0, 0, 0, 0);     // consumed
  else if (!all && c) for (int k = 0; k < 16 && e0 + k < n; k++) dirty[e0 + k] = 0;
}
__global__ void __launch_bounds__(256) pack_records_kernel(const uint8_t* __restrict__ obs, uint8_t* __restrict__ staging, long long cap) {
  const int total = *reinterpret_cast<const int*>(staging);
  const uint32_t* idx = reinterpret_cast<const uint32_t*>(staging + 16);
  uint8_t* rec_out = staging + dirty_rec_offset(cap);
  const int lane = threadIdx.x & 31, sub = lane / DELTA_LANES, part = lane - sub * DELTA_LANES;
  const long long warp_gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long warp_cnt = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long k0 = warp_gid * DELTA_PER_WARP; k0 < total; k0 += warp_cnt * DELTA_PER_WARP) {
    const long long k = k0 + sub;
    if (sub < DELTA_PER_WARP && k < total && k < cap) {
      const long long e = idx[k] & 0x7FFFFFFFu;
      reinterpret_cast<uint4*>(rec_out + k * BGYM_OBS_DELTA_BYTES)[part] = __ldcs(reinterpret_cast<const uint4*>(obs + e * BGYM_OBS_BYTES) + part);
    }
  }
}
// staged record k -> the mirror (BGYM_MIRROR_CORE_BYTES = 128: chunks 0..5, 8, 9 of the record at core[e]; BGYM_MIRROR_SHOP_BYTES
// = 32: chunks 6, 7 at shop[e]).  A core record is one aligned 128-byte line WRITTEN BY ONE STORE INSTRUCTION (8 lanes x 16 B,
// four records per warp): zero-copy stores into pinned host memory then run at the link's full rate (measured 52 GB/s against
// 35-42 GB/s for 176-byte records at a 176-byte stride, tools/exp/zc_probe.cu; the same line written by two instructions —
// 96 + 32 bytes — fell to 25 GB/s).
__global__ void __launch_bounds__(256) scatter_dirty_kernel(const uint8_t* __restrict__ staging, long long cap, uint8_t* __restrict__ core,
                                                            uint8_t* __restrict__ shop) {
  const int total = *reinterpret_cast<const int*>(staging);
  const uint32_t* idx = reinterpret_cast<const uint32_t*>(staging + 16);
  const uint8_t* rec = staging + dirty_rec_offset(cap);
  const int lane = threadIdx.x & 31, sub = lane >> 3, part = lane & 7;
  const int chunk = part < 6 ? part : part + 2;
  const long long warp_gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long warp_cnt = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long k0 = warp_gid * 4; k0 < total; k0 += warp_cnt * 4) {
    const long long k = k0 + sub;
    if (k < total && k < cap) {
      const uint32_t w = __ldg(idx + k);
      const long long e = w & 0x7FFFFFFFu;
      const uint4* src = reinterpret_cast<const uint4*>(rec + k * BGYM_OBS_DELTA_BYTES);
      reinterpret_cast<uint4*>(core + e * BGYM_MIRROR_CORE_BYTES)[part] = __ldcs(src + chunk);
      if ((w & 0x80000000u) && part < 2) reinterpret_cast<uint4*>(shop + e * BGYM_MIRROR_SHOP_BYTES)[part] = __ldcs(src + 6 + part);
    }
  }
}

}  // namespace bgym

// =================================================================================================
// C-ABI
// =================================================================================================
using namespace bgym;

static thread_local char g_err[256] = "";
static int set_err(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}
static int cuda_rc(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof g_err, "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

// Per-device one-time setup (kernel attributes are per device) under a mutex: the entry points may be called
// from several host threads / for several devices of one process.  g_sm_count is the calling thread's device.
constexpr int BGYM_MAX_DEVICES = 64;
static std::mutex g_mu;
static bool g_dev_ready[BGYM_MAX_DEVICES];
static int g_dev_sms[BGYM_MAX_DEVICES];
static thread_local int g_sm_count = 0;
static int ensure_device_setup() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_rc(e, "cudaGetDevice");
  if (dev < 0 || dev >= BGYM_MAX_DEVICES) return set_err(BGYM_E_ARG, "device ordinal out of range");
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_dev_ready[dev]) { g_sm_count = g_dev_sms[dev]; return 0; }
  e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_rc(e, "cudaDeviceGetAttribute");
#define BGYM_SET_SMEM(kernel, bytes)                                                                 \
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);             \
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(" #kernel ")");
  BGYM_SET_SMEM(env_reset_kernel, RESET_CTA_SMEM)
  BGYM_SET_SMEM(policy_first_layer_kernel, FL_SMEM)
#undef BGYM_SET_SMEM
  g_dev_sms[dev] = g_sm_count;
  g_dev_ready[dev] = true;
  return 0;
}

// slabs of at most this many envs take the one-launch step (BGYM_SMALL_N in the environment, or bgym_set_option)
static long long g_small_n = -1;
static long long small_slab_threshold() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_small_n < 0) { const char* v = getenv("BGYM_SMALL_N"); g_small_n = v ? atoll(v) : (long long)BGYM_SMALL_N_DEFAULT; }
  return g_small_n;
}

static bool misaligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

// scratch of the partitioned step (deferred-env lists, their counters, the per-env continuation word), one set per
// (device, stream) in use; bgym_release_stream() gives a set back
// default launch plan of level 1 (see bgym_step): streams by list, launch order
#define BGYM_L1_PLAN_STREAMS "0121230"
#define BGYM_L1_PLAN_STREAMS_FUSED "0123456"
#define BGYM_L1_PLAN_ORDER "2105463"
#define BGYM_L1_PLAN_ORDER_FUSED "2105463"
constexpr int PART_SIDE_STREAMS = 6;   // the seven level-1 list kernels run concurrently: launch stream + six forked ones
struct PartScratch { bool used; int dev; void* stream; long long cap; int* lists; int* counters; uint16_t* aux;

                     cudaStream_t side[PART_SIDE_STREAMS]; cudaEvent_t ev_fork, ev_fork2, ev_side[PART_SIDE_STREAMS]; bool streams_ok; };
constexpr int BGYM_MAX_SCRATCH = 64;
static PartScratch g_scratch[BGYM_MAX_SCRATCH];
static int get_part_scratch(long long n, void* stream, PartScratch** out) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_mu);
  PartScratch* sc = nullptr;
  for (int i = 0; i < BGYM_MAX_SCRATCH; i++)
    if (g_scratch[i].used && g_scratch[i].dev == dev && g_scratch[i].stream == stream) sc = &g_scratch[i];
  if (!sc) {
    for (int i = 0; i < BGYM_MAX_SCRATCH && !sc; i++) if (!g_scratch[i].used) sc = &g_scratch[i];
    if (!sc) return set_err(BGYM_E_ARG, "bgym_step: too many (device, stream) pairs in use (bgym_release_stream frees one)");
    sc->used = true; sc->dev = dev; sc->stream = stream; sc->cap = 0; sc->lists = nullptr; sc->counters = nullptr; sc->aux = nullptr;
    sc->streams_ok = cudaEventCreateWithFlags(&sc->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&sc->ev_fork2, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < PART_SIDE_STREAMS; i++)
      sc->streams_ok = sc->streams_ok && cudaStreamCreateWithFlags(&sc->side[i], cudaStreamNonBlocking) == cudaSuccess &&
                       cudaEventCreateWithFlags(&sc->ev_side[i], cudaEventDisableTiming) == cudaSuccess;
  }
  if (sc->cap < n) {
    if (sc->lists) cudaFree(sc->lists);
    const size_t ints = (size_t)N_LISTS * (size_t)n + (size_t)N_LISTS * PART_CTR_STRIDE;
    cudaError_t e = cudaMalloc(&sc->lists, ints * sizeof(int) + (size_t)n * sizeof(uint16_t));
    if (e != cudaSuccess) { sc->cap = 0; sc->lists = nullptr; return cuda_rc(e, "cudaMalloc(step scratch)"); }
    sc->cap = n;
    sc->counters = sc->lists + (size_t)N_LISTS * (size_t)n;
    sc->aux = reinterpret_cast<uint16_t*>(sc->counters + N_LISTS * PART_CTR_STRIDE);
  }
  *out = sc;
  return 0;
}

// a launch-plan override (bgym_step): seven digits 0..6, as an order each exactly once; anything else -> nullptr (ignored)
static const char* valid_plan(const char* v, bool permutation) {
  if (!v || strlen(v) != (size_t)N_LISTS_L1) return nullptr;
  int seen = 0;
  for (int i = 0; i < N_LISTS_L1; i++) {
    const int k = v[i] - '0';
    if (k < 0 || k > PART_SIDE_STREAMS) return nullptr;
    seen |= 1 << k;
  }
  if (permutation && seen != (1 << N_LISTS_L1) - 1) return nullptr;
  return v;
}

static int tile_grid(long long n, int warps, int ctas_per_sm) {
  long long tiles = (n + 31) / 32;
  long long ctas = (tiles + warps - 1) / warps;
  long long cap = (long long)g_sm_count * ctas_per_sm;   // persistent: resident CTAs only
  return (int)(ctas < cap ? ctas : cap);
}

extern "C" {

int bgym_abi_version(void) { return BGYM_ABI_VERSION; }
int bgym_set_option(int option, int64_t value) {
  if (option == BGYM_OPT_SMALL_SLAB) {
    if (value < 0) return set_err(BGYM_E_ARG, "bgym_set_option: BGYM_OPT_SMALL_SLAB needs a value >= 0");
    std::lock_guard<std::mutex> lock(g_mu);
    g_small_n = value;
    return 0;
  }
  return set_err(BGYM_E_ARG, "bgym_set_option: unknown option");
}
const char* bgym_last_error(void) { return g_err; }
int bgym_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int bgym_reset(BgymHot* hot, BgymTog* tog, BgymCold* cold, BgymObs* obs, BgymSel* sel, const uint8_t* reset_mask,
               const uint32_t* seeds, const uint8_t* decks52, int64_t n, int flags, void* stream) {
  if (n < 0 || !hot || !tog || !cold || !seeds) return set_err(BGYM_E_ARG, "bgym_reset: null hot/tog/cold/seeds or negative n");
  if ((!obs || !sel) && !(flags & BGYM_FLAG_NO_OBS)) return set_err(BGYM_E_ARG, "bgym_reset: obs / sel is NULL without BGYM_FLAG_NO_OBS");
  if (misaligned(hot, 16) || misaligned(tog, 16) || misaligned(cold, 16) || misaligned(obs, 16) || misaligned(sel, 16))
    return set_err(BGYM_E_ALIGN, "bgym_reset: hot/tog/cold/obs/sel must be 16-byte aligned");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.hot = reinterpret_cast<uint8_t*>(hot); a.tog = reinterpret_cast<uint8_t*>(tog); a.cold = reinterpret_cast<uint8_t*>(cold);
  a.obs = reinterpret_cast<uint8_t*>(obs); a.sel = reinterpret_cast<uint8_t*>(sel);
  if (!obs || !sel) { a.obs = nullptr; a.sel = nullptr; }
  a.reset_mask = reset_mask; a.seeds = seeds; a.decks52 = decks52; a.n = n; a.flags = flags;
  env_reset_kernel<<<tile_grid(n, RESET_WARPS, 3), RESET_WARPS * 32, RESET_CTA_SMEM, (cudaStream_t)stream>>>(a);
  return cuda_rc(cudaGetLastError(), "bgym_reset launch");
}

int bgym_step(BgymHot* hot, BgymTog* tog, BgymCold* cold, int32_t* actions, const BgymDraws* draws, BgymObs* obs, BgymSel* sel,
              uint8_t* obs_dirty, double* reward, uint8_t* terminated, uint8_t* truncated, BgymInfo* info,
              int64_t n, int flags, void* stream) {
  if (n < 0 || !hot || !tog || !cold || !reward || !terminated)
    return set_err(BGYM_E_ARG, "bgym_step: null hot/tog/cold/reward/terminated or negative n");
  if (!actions) return set_err(BGYM_E_ARG, "bgym_step: actions is NULL");
  if ((!obs || !sel) && !(flags & BGYM_FLAG_NO_OBS)) return set_err(BGYM_E_ARG, "bgym_step: obs / sel is NULL without BGYM_FLAG_NO_OBS");
  if (misaligned(hot, 16) || misaligned(tog, 16) || misaligned(cold, 16) || misaligned(obs, 16) || misaligned(sel, 16) ||
      misaligned(info, 16) || misaligned(draws, 8))
    return set_err(BGYM_E_ALIGN, "bgym_step: hot/tog/cold/obs/sel/info must be 16-byte aligned");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  PartScratch* sc = nullptr;
  rc = get_part_scratch(n, stream, &sc);
  if (rc) return rc;
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.hot = reinterpret_cast<uint8_t*>(hot); a.tog = reinterpret_cast<uint8_t*>(tog); a.cold = reinterpret_cast<uint8_t*>(cold);
  a.actions = actions;
  a.actions_out = (flags & BGYM_FLAG_RANDOM_POLICY) ? actions : nullptr;
  a.draws = draws; a.obs = reinterpret_cast<uint8_t*>(obs); a.sel = reinterpret_cast<uint8_t*>(sel);
  if (!obs || !sel) { a.obs = nullptr; a.sel = nullptr; }
  a.obs_dirty = a.obs ? obs_dirty : nullptr;

  a.reward = reward; a.terminated = terminated; a.truncated = truncated; a.info = info;
  a.n = n; a.flags = flags;
  a.part_lists = sc->lists; a.part_counters = sc->counters; a.part_aux = sc->aux; a.part_cap = sc->cap;
  // BGYM_STEP_TIMING (diagnostic, single-threaded use): synchronising timing of the launch set, printed every 64 calls.
  //   1 = phases (main / level-1 lists concurrently / level-2 lists);  2 = every kernel on its own, launched serially
  static const int timing = getenv("BGYM_STEP_TIMING") ? atoi(getenv("BGYM_STEP_TIMING")) : 0;
  constexpr int NEV = N_LISTS + 2;
  static cudaEvent_t tev[NEV];
  static double tsum[NEV] = {0};
  static int tcalls = 0;
  if (timing && tcalls == 0 && tsum[0] == 0) for (int i = 0; i < NEV; i++) cudaEventCreate(&tev[i]);
  if (timing) cudaEventRecord(tev[0], s);
  // small slabs: one launch (bgym_step_part.cuh, env_step_small_kernel); BGYM_SMALL_N overrides the threshold
  if (n <= small_slab_threshold() && !timing) {
    env_step_small_kernel<<<tile_grid(n, GATHER_WARPS, GATHER_CTAS_PER_SM), GATHER_WARPS * 32, GATHER_CTA_SMEM, s>>>(a);
    return cuda_rc(cudaGetLastError(), "bgym_step launch");
  }
  cudaError_t e = cudaMemsetAsync(sc->counters, 0, N_LISTS * PART_CTR_STRIDE * sizeof(int), s);
  if (e != cudaSuccess) return cuda_rc(e, "cudaMemsetAsync(step counters)");
  env_step_main_kernel<<<tile_grid(n, MAIN_WARPS, MAIN_CTAS_PER_SM), MAIN_WARPS * 32, 0, s>>>(a);
  if (timing) cudaEventRecord(tev[1], s);
  // The list lengths live on the device: every list kernel is launched with a resident-size grid, idle warps exit at
  // once.  Level 1 = one kernel per list the main pass filled, run concurrently on forked streams (they touch disjoint
  // envs); level 2 = the round advances and resets level 1 handed on, again concurrently.  (Chaining ADVANCE behind
  // PLAY and RESET behind PLAY / CONS / MISC only, so that level 2 overlaps the rest of level 1, measured 4 % slower:
  // the passes are bound by resident warps, not by idle SMs.)
  const int gt = GATHER_WARPS * 32;
  auto lgrid = [&](int list) { int g = tile_grid(n, GATHER_WARPS, list_ctas(list)); return g < 1 ? 1 : g; };
  // BGYM_LEVEL_KERNELS (bit 0: level 2, bit 1: level 1): one kernel per level on the launch stream
  // (env_step_level_kernel) instead of one per list on forked streams
  // Measured (tools/exp/level_kernels.sh): level 2 as one kernel 0.1841 -> 0.1830 ms per step (one launch instead of a
  // fork, two launches and a join) — the default; level 1 as one kernel with out-of-line bodies and claimed tiles 0.197 ms
  // against 0.177 (inlined bodies: 0.218 ms, 255 registers and 2 KB of spills).
  static const int level_kernels = getenv("BGYM_LEVEL_KERNELS") ? atoi(getenv("BGYM_LEVEL_KERNELS")) : 1;
  if ((level_kernels & 2) && !timing) {
    env_step_level_kernel<1><<<g_sm_count * level_ctas(1), 32, 0, s>>>(a);
    env_step_level_kernel<2><<<g_sm_count * level_ctas(2), 32, 32 * BGYM_COLD_BYTES, s>>>(a);
    return cuda_rc(cudaGetLastError(), "bgym_step launch");
  }
  static const bool serial = getenv("BGYM_SERIAL_GATHER") != nullptr;
  const bool fork = sc->streams_ok && !serial && timing != 2;
  cudaStream_t ls[N_LISTS];
  for (int l = 0; l < N_LISTS; l++) ls[l] = s;
  // Level-1 launch plan: which stream each list kernel goes to (0 = the launch stream, k = forked stream k - 1) and the
  // launch order.  Kernels on one stream run one after the other, kernels on different streams concurrently — in an order
  // the block scheduler picks, and the level's duration moves by +-5 % with it (profiles/r02_step_experiments.md).  The plan
  // puts the lists with long tiles at the head of the streams and chains the lists with short tiles behind them, so that
  // the level ends on short tiles.  BGYM_L1_STREAMS / BGYM_L1_ORDER (7 digits each, indexed / listing PLAY CONS GEN MISC
  // DISCARD SHOP BLIND = 0..6) override it; with the fused policy every env outside PLAY phase is in MISC.
  // (an override that is not seven digits — streams 0..6, order a permutation of 0..6 — is ignored: a list that is never
  // launched would leave its envs unstepped)
  static const char* plan_streams_env = valid_plan(getenv("BGYM_L1_STREAMS"), false);
  static const char* plan_order_env = valid_plan(getenv("BGYM_L1_ORDER"), true);
  static const char* plan_streams_f_env = valid_plan(getenv("BGYM_L1_STREAMS_FUSED"), false);
  static const char* plan_order_f_env = valid_plan(getenv("BGYM_L1_ORDER_FUSED"), true);
  const bool fusedp = (flags & BGYM_FLAG_RANDOM_POLICY) != 0;
  const char* plan_streams = fusedp ? (plan_streams_f_env ? plan_streams_f_env : BGYM_L1_PLAN_STREAMS_FUSED)
                                    : (plan_streams_env ? plan_streams_env : BGYM_L1_PLAN_STREAMS);
  const char* plan_order = fusedp ? (plan_order_f_env ? plan_order_f_env : BGYM_L1_PLAN_ORDER_FUSED)
                                  : (plan_order_env ? plan_order_env : BGYM_L1_PLAN_ORDER);
  int n_side = 0;
  if (fork) {
    for (int l = 0; l < N_LISTS_L1; l++) {
      int k = plan_streams[l] - '0';
      if (k < 0 || k > PART_SIDE_STREAMS) k = 0;
      ls[l] = k == 0 ? s : sc->side[k - 1];
      if (k > n_side) n_side = k;
    }
    cudaEventRecord(sc->ev_fork, s);
    for (int i = 0; i < n_side; i++) cudaStreamWaitEvent(sc->side[i], sc->ev_fork, 0);
  }
  for (int q = 0; q < N_LISTS_L1; q++) {
    const int l = fork ? plan_order[q] - '0' : q;
    switch (l) {
#define BGYM_LAUNCH_LIST(L) case L: env_step_list_kernel<L><<<lgrid(L), gt, list_smem_bytes(L), ls[L]>>>(a); if (timing == 2) cudaEventRecord(tev[2 + L], s); break;
      BGYM_LAUNCH_LIST(L_PLAY) BGYM_LAUNCH_LIST(L_CONS) BGYM_LAUNCH_LIST(L_GEN) BGYM_LAUNCH_LIST(L_MISC)
      BGYM_LAUNCH_LIST(L_DISCARD) BGYM_LAUNCH_LIST(L_SHOP) BGYM_LAUNCH_LIST(L_BLIND)
#undef BGYM_LAUNCH_LIST
      default: break;
    }
  }
  if (fork)
    for (int i = 0; i < n_side; i++) { cudaEventRecord(sc->ev_side[i], sc->side[i]); cudaStreamWaitEvent(s, sc->ev_side[i], 0); }
  if (timing == 1) cudaEventRecord(tev[2], s);
  if ((level_kernels & 1) && timing != 2) {
    env_step_level_kernel<2><<<g_sm_count * level_ctas(2), 32, 32 * BGYM_COLD_BYTES, s>>>(a);
  } else {
    if (fork) { cudaEventRecord(sc->ev_fork2, s); cudaStreamWaitEvent(sc->side[0], sc->ev_fork2, 0); ls[L_RESET] = sc->side[0]; }
    env_step_list_kernel<L_ADVANCE><<<lgrid(L_ADVANCE), gt, list_smem_bytes(L_ADVANCE), s>>>(a); if (timing == 2) cudaEventRecord(tev[2 + L_ADVANCE], s);
    env_step_list_kernel<L_RESET><<<lgrid(L_RESET), gt, list_smem_bytes(L_RESET), ls[L_RESET]>>>(a); if (timing == 2) cudaEventRecord(tev[2 + L_RESET], s);
    if (fork) { cudaEventRecord(sc->ev_side[0], sc->side[0]); cudaStreamWaitEvent(s, sc->ev_side[0], 0); }
  }
  if (timing == 1) cudaEventRecord(tev[3], s);
#ifdef BGYM_TILE_CLOCK
  {
    static int dbg_calls = 0;
    if (++dbg_calls % 64 == 0) {
      cudaStreamSynchronize(s);
      static int hc[N_LISTS * PART_CTR_STRIDE];
      cudaMemcpy(hc, sc->counters, sizeof hc, cudaMemcpyDeviceToHost);
      static const char* lnames[N_LISTS] = {"PLAY", "CONS", "GEN", "MISC", "DISCARD", "SHOP", "BLIND", "ADVANCE", "RESET"};
      unsigned long long t_ref = ~0ull;
      for (int l = 0; l < N_LISTS; l++) {
        const unsigned long long* d = reinterpret_cast<const unsigned long long*>(hc + l * PART_CTR_STRIDE + 8);
        if (d[4] && ~d[0] < t_ref) t_ref = ~d[0];
      }
      fprintf(stderr, "[bgym tiles]");
      for (int l = 0; l < N_LISTS; l++) {
        const unsigned long long* d = reinterpret_cast<const unsigned long long*>(hc + l * PART_CTR_STRIDE + 8);
        if (!d[4]) continue;
        fprintf(stderr, " %s n=%llu start %.1f end %.1f mean %.1f max %.1f |", lnames[l], d[4], (~d[0] - t_ref) * 1e-3, (d[1] - t_ref) * 1e-3,
                d[2] * 1e-3 / d[4], d[3] * 1e-3);
      }
      fprintf(stderr, " (us)\n");
    }
  }
#endif
  if (timing == 1) {
    cudaEventSynchronize(tev[3]);
    for (int i = 0; i < 3; i++) { float m = 0; cudaEventElapsedTime(&m, tev[i], tev[i + 1]); tsum[i] += m; }
    if (++tcalls % 64 == 0) {
      fprintf(stderr, "[bgym timing] main %.1f us, level-1 lists %.1f us, level-2 lists %.1f us (mean of 64 steps)\n",
              tsum[0] / 64 * 1e3, tsum[1] / 64 * 1e3, tsum[2] / 64 * 1e3);
      for (int i = 0; i < 3; i++) tsum[i] = 1e-30;
    }
  } else if (timing == 2) {
    cudaEventSynchronize(tev[2 + L_RESET]);
    if (++tcalls % 64 == 0) {
      static const int order[N_LISTS] = {L_PLAY, L_CONS, L_GEN, L_MISC, L_DISCARD, L_SHOP, L_BLIND, L_ADVANCE, L_RESET};
      static const char* names[N_LISTS] = {"PLAY", "CONS", "GEN", "MISC", "DISCARD", "SHOP", "BLIND", "ADVANCE", "RESET"};
      fprintf(stderr, "[bgym timing, serial launches, us]");
      float m = 0;
      cudaEventElapsedTime(&m, tev[0], tev[1]);
      fprintf(stderr, " main %.1f", m * 1e3);
      int prev = 1;
      for (int k = 0; k < N_LISTS; k++) {
        cudaEventElapsedTime(&m, tev[prev], tev[2 + order[k]]);
        fprintf(stderr, " %s %.1f", names[order[k]], m * 1e3);
        prev = 2 + order[k];
      }
      fprintf(stderr, " (last step of 64)\n");
    }
  }
  return cuda_rc(cudaGetLastError(), "bgym_step launch");
}

// gives back the step scratch held for `stream` on the current device (the lists are sized for the largest slab the
// stream has stepped).  Call it after the stream's last step has finished, e.g. before destroying the stream.
int bgym_release_stream(void* stream) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_err(BGYM_E_NODEV, "bgym_release_stream: no CUDA device");
  std::lock_guard<std::mutex> lock(g_mu);
  for (int i = 0; i < BGYM_MAX_SCRATCH; i++)
    if (g_scratch[i].used && g_scratch[i].dev == dev && g_scratch[i].stream == stream) {
      if (g_scratch[i].lists) cudaFree(g_scratch[i].lists);
      cudaEventDestroy(g_scratch[i].ev_fork); cudaEventDestroy(g_scratch[i].ev_fork2);
      for (int k = 0; k < PART_SIDE_STREAMS; k++) { cudaStreamDestroy(g_scratch[i].side[k]); cudaEventDestroy(g_scratch[i].ev_side[k]); }
      g_scratch[i] = PartScratch{};
    }
  return 0;
}

int bgym_action_mask(const BgymHot* hot, const BgymTog* tog, const BgymCold* cold, uint64_t* mask, int64_t n, void* stream) {
  if (n < 0 || !hot || !tog || !cold || !mask) return set_err(BGYM_E_ARG, "bgym_action_mask: bad arguments");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  int grid = (int)((n + 255) / 256 < (long long)g_sm_count * 8 ? (n + 255) / 256 : (long long)g_sm_count * 8);
  action_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(hot), reinterpret_cast<const uint8_t*>(tog),
                                                             reinterpret_cast<const uint8_t*>(cold), mask, n);
  return cuda_rc(cudaGetLastError(), "bgym_action_mask launch");
}

static int sample_grid(int64_t n) {
  // one thread per env up to 64 resident-size waves
  return (int)((n + 255) / 256 < (long long)g_sm_count * 512 ? (n + 255) / 256 : (long long)g_sm_count * 512);
}
static int check_mask_words(const uint64_t* mask_words, int64_t mask_stride, const char* who) {
  if (!mask_words || mask_stride < 8 || (mask_stride & 7) || misaligned(mask_words, 8)) {
    snprintf(g_err, sizeof g_err, "%s: mask_words must be 8-byte aligned and mask_stride a multiple of 8 (>= 8)", who);
    return BGYM_E_ARG;
  }
  return 0;
}

int bgym_sample_actions(const uint64_t* mask_words, int64_t mask_stride, int32_t* actions, uint32_t seed, uint64_t step,
                        int64_t n, void* stream) {
  if (n < 0 || !actions) return set_err(BGYM_E_ARG, "bgym_sample_actions: bad arguments");
  int rc = check_mask_words(mask_words, mask_stride, "bgym_sample_actions");
  if (rc) return rc;
  if (n == 0) return 0;
  rc = ensure_device_setup();
  if (rc) return rc;
  sample_actions_kernel<<<sample_grid(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(mask_words), mask_stride, actions, seed, step, nullptr, n);
  return cuda_rc(cudaGetLastError(), "bgym_sample_actions launch");
}

int bgym_sample_actions_ctr(const uint64_t* mask_words, int64_t mask_stride, int32_t* actions, uint32_t seed,
                            uint64_t* step_counter, int64_t n, void* stream) {
  if (n < 0 || !actions || !step_counter) return set_err(BGYM_E_ARG, "bgym_sample_actions_ctr: bad arguments");
  int rc = check_mask_words(mask_words, mask_stride, "bgym_sample_actions_ctr");
  if (rc) return rc;
  if (n == 0) return 0;
  rc = ensure_device_setup();
  if (rc) return rc;
  sample_actions_kernel<<<sample_grid(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(mask_words), mask_stride, actions, seed, 0ull,
                                                                          reinterpret_cast<const unsigned long long*>(step_counter), n);
  bump_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(step_counter));
  return cuda_rc(cudaGetLastError(), "bgym_sample_actions_ctr launch");
}

int bgym_sync_state(BgymHot* hot, BgymTog* tog, int64_t n, int direction, void* stream) {
  if (n < 0 || !hot || !tog || (direction != BGYM_SYNC_TO_RECORDS && direction != BGYM_SYNC_FROM_RECORDS))
    return set_err(BGYM_E_ARG, "bgym_sync_state: bad arguments");
  if (misaligned(hot, 16) || misaligned(tog, 16)) return set_err(BGYM_E_ALIGN, "bgym_sync_state: hot/tog must be 16-byte aligned");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  sync_state_kernel<<<sample_grid(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint8_t*>(hot), reinterpret_cast<uint8_t*>(tog), n, direction);
  return cuda_rc(cudaGetLastError(), "bgym_sync_state launch");
}

int bgym_sync_obs(BgymObs* obs, BgymSel* sel, int64_t n, int direction, void* stream) {
  if (n < 0 || !obs || !sel || (direction != BGYM_SYNC_TO_RECORDS && direction != BGYM_SYNC_FROM_RECORDS))
    return set_err(BGYM_E_ARG, "bgym_sync_obs: bad arguments");
  if (misaligned(obs, 16) || misaligned(sel, 16)) return set_err(BGYM_E_ALIGN, "bgym_sync_obs: obs/sel must be 16-byte aligned");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  sync_obs_kernel<<<sample_grid(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint8_t*>(obs), reinterpret_cast<uint8_t*>(sel), n, direction);
  return cuda_rc(cudaGetLastError(), "bgym_sync_obs launch");
}

int bgym_pack_dirty_obs(const BgymObs* obs, uint8_t* obs_dirty, void* staging, void* scratch, int64_t cap, int64_t n, void* stream) {
  if (!obs || !staging || !scratch || cap <= 0 || n <= 0) return set_err(BGYM_E_ARG, "bgym_pack_dirty_obs: bad arguments");
  if (misaligned(obs, 16) || misaligned(staging, 16) || misaligned(obs_dirty, 16) || misaligned(scratch, 4))
    return set_err(BGYM_E_ALIGN, "bgym_pack_dirty_obs: obs/staging/obs_dirty must be 16-byte aligned");
  if (cap < n) return set_err(BGYM_E_ARG, "bgym_pack_dirty_obs: cap must be at least n");
  int rc = ensure_device_setup();
  if (rc) return rc;
  const int all = obs_dirty ? 0 : 1;
  const long long blocks = (n + PACK_ENVS - 1) / PACK_ENVS;
  if (blocks > 0x7fffffffLL) return set_err(BGYM_E_ARG, "bgym_pack_dirty_obs: n too large");
  cudaStream_t s = (cudaStream_t)stream;
  pack_count_kernel<<<(int)blocks, PACK_THREADS, 0, s>>>(obs_dirty, reinterpret_cast<int*>(scratch), n, all);
  pack_index_kernel<<<(int)blocks, PACK_THREADS, 0, s>>>(obs_dirty, reinterpret_cast<const int*>(scratch), reinterpret_cast<uint8_t*>(staging), cap, n, all);
  pack_records_kernel<<<g_sm_count * 8, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(obs), reinterpret_cast<uint8_t*>(staging), cap);
  return cuda_rc(cudaGetLastError(), "bgym_pack_dirty_obs launch");
}

int bgym_scatter_dirty_obs(const void* staging, int64_t cap, void* mirror_core, void* mirror_shop, void* stream) {
  if (!staging || !mirror_core || !mirror_shop || cap <= 0) return set_err(BGYM_E_ARG, "bgym_scatter_dirty_obs: bad arguments");
  if (misaligned(mirror_core, 128) || misaligned(mirror_shop, 32) || misaligned(staging, 16))
    return set_err(BGYM_E_ALIGN, "bgym_scatter_dirty_obs: mirror_core needs 128-byte, mirror_shop 32-byte, staging 16-byte alignment");
  int rc = ensure_device_setup();
  if (rc) return rc;
  scatter_dirty_kernel<<<g_sm_count * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(staging), cap,
                                                                        reinterpret_cast<uint8_t*>(mirror_core), reinterpret_cast<uint8_t*>(mirror_shop));
  return cuda_rc(cudaGetLastError(), "bgym_scatter_dirty_obs launch");
}

int bgym_score_hands(const uint8_t* cards8, const uint16_t* mods8, const uint8_t* n_cards,
                     const uint8_t* jokers8, const uint8_t* levels12, const BgymScoreCtx* ctx,
                     uint8_t* hand_type, int32_t* chips, int32_t* mult, double* x_mult,
                     int64_t* score, int32_t* money, uint32_t seed, int64_t n, int flags, void* stream) {
  if (n < 0 || !cards8 || !hand_type || !chips || !mult || !score) return set_err(BGYM_E_ARG, "bgym_score_hands: bad arguments");
  if (misaligned(cards8, 8) || misaligned(mods8, 16) || misaligned(jokers8, 8) || misaligned(ctx, 16))
    return set_err(BGYM_E_ALIGN, "bgym_score_hands: cards8/jokers8 need 8-byte, mods8/ctx 16-byte alignment");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  ScoreArgs a;
  a.cards8 = cards8; a.mods8 = mods8; a.n_cards = n_cards; a.jokers8 = jokers8; a.levels12 = levels12; a.ctx = ctx;
  a.hand_type = hand_type; a.chips = chips; a.mult = mult; a.x_mult = x_mult; a.score = reinterpret_cast<long long*>(score);
  a.money = money; a.seed = seed; a.n = n; a.flags = flags;
  long long blocks = (n + 255) / 256;
  long long cap = (long long)g_sm_count * 16;
  if (!mods8 && !n_cards && !jokers8 && !levels12 && !ctx && !(flags & BGYM_SCORE_RULES))
    score_hands5_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(a);
  else
    score_hands_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(a);
  return cuda_rc(cudaGetLastError(), "bgym_score_hands launch");
}

int bgym_episode_stats(const double* reward, const uint8_t* terminated, double* ret_acc, uint32_t* len_acc,
                       double* stats, int64_t n, void* stream) {
  if (n < 0 || !reward || !terminated || !ret_acc || !len_acc || !stats) return set_err(BGYM_E_ARG, "bgym_episode_stats: bad arguments");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  int grid = (int)((n + 255) / 256 < (long long)g_sm_count * 8 ? (n + 255) / 256 : (long long)g_sm_count * 8);
  episode_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reward, terminated, ret_acc, len_acc, stats, n);
  return cuda_rc(cudaGetLastError(), "bgym_episode_stats launch");
}

// ---- rollout collection kernels (bgym_rollout.cuh) ------------------------------------------------
int bgym_featurize(const BgymObs* obs, void* features, int64_t n, int dtype, void* stream) {
  if (n < 0 || !obs || !features) return set_err(BGYM_E_ARG, "bgym_featurize: bad arguments");
  if (dtype != BGYM_DT_F32 && dtype != BGYM_DT_BF16) return set_err(BGYM_E_ARG, "bgym_featurize: dtype must be BGYM_DT_F32 or BGYM_DT_BF16");
  if (misaligned(obs, 16) || misaligned(features, 16)) return set_err(BGYM_E_ALIGN, "bgym_featurize: obs/features must be 16-byte aligned");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  long long blocks = (n + FEAT_ENVS_PER_CTA - 1) / FEAT_ENVS_PER_CTA;
  long long cap = (long long)g_sm_count * 64;
  int grid = (int)(blocks < cap ? blocks : cap);
  const int ft = FEAT_ENVS_PER_CTA * FEAT_CHUNKS;
  if (dtype == BGYM_DT_F32)
    featurize_kernel<float><<<grid, ft, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(obs), reinterpret_cast<float*>(features), n);
  else
    featurize_kernel<__nv_bfloat16><<<grid, ft, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(obs), reinterpret_cast<__nv_bfloat16*>(features), n);
  return cuda_rc(cudaGetLastError(), "bgym_featurize launch");
}

int bgym_policy_first_layer(const BgymObs* obs, const void* wt_hand, const void* wt_joker, const void* wt_game,
                            const float* bias, void* out, int64_t n, void* stream) {
  if (n < 0 || !obs || !wt_hand || !wt_joker || !wt_game || !bias || !out) return set_err(BGYM_E_ARG, "bgym_policy_first_layer: bad arguments");
  if (misaligned(obs, 16) || misaligned(wt_hand, 16) || misaligned(wt_joker, 16) || misaligned(wt_game, 16) || misaligned(out, 16))
    return set_err(BGYM_E_ALIGN, "bgym_policy_first_layer: obs / weights / out must be 16-byte aligned");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  long long ctas = (n + FL_WARPS - 1) / FL_WARPS;
  if (ctas > g_sm_count) ctas = g_sm_count;                  // persistent: one CTA per SM holds the weights in shared memory
  policy_first_layer_kernel<<<(int)ctas, FL_WARPS * 32, FL_SMEM, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(obs), reinterpret_cast<const __nv_bfloat16*>(wt_hand), reinterpret_cast<const __nv_bfloat16*>(wt_joker),
      reinterpret_cast<const __nv_bfloat16*>(wt_game), bias, reinterpret_cast<__nv_bfloat16*>(out), n);
  return cuda_rc(cudaGetLastError(), "bgym_policy_first_layer launch");
}

int bgym_masked_sample(const void* logits, int dtype, const uint64_t* mask_words, int64_t mask_stride, const float* uniforms,
                       uint32_t seed, uint64_t step, int64_t env_offset,
                       int32_t* actions, float* logp, float* entropy, int64_t n, void* stream) {
  if (n < 0 || !logits || !actions || !logp) return set_err(BGYM_E_ARG, "bgym_masked_sample: bad arguments");
  if (dtype != BGYM_DT_F32 && dtype != BGYM_DT_BF16) return set_err(BGYM_E_ARG, "bgym_masked_sample: dtype must be BGYM_DT_F32 or BGYM_DT_BF16");
  int rc = check_mask_words(mask_words, mask_stride, "bgym_masked_sample");
  if (rc) return rc;
  if (misaligned(logits, dtype == BGYM_DT_F32 ? 16 : 8))
    return set_err(BGYM_E_ALIGN, "bgym_masked_sample: logits need 16-byte (f32) / 8-byte (bf16) alignment");
  if (n == 0) return 0;
  rc = ensure_device_setup();
  if (rc) return rc;
  long long blocks = (n + 127) / 128;
  if (blocks > 0x7fffffffLL) return set_err(BGYM_E_ARG, "bgym_masked_sample: n too large for one launch");
  if (dtype == BGYM_DT_F32)
    masked_sample_kernel<float><<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(logits), reinterpret_cast<const uint8_t*>(mask_words), mask_stride,
        uniforms, seed, step, env_offset, actions, logp, entropy, n);
  else
    masked_sample_kernel<__nv_bfloat16><<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), reinterpret_cast<const uint8_t*>(mask_words), mask_stride,
        uniforms, seed, step, env_offset, actions, logp, entropy, n);
  return cuda_rc(cudaGetLastError(), "bgym_masked_sample launch");
}

int bgym_gae(const float* rewards, const float* values, const uint8_t* dones, float gamma, float lam,
             float* advantages, float* returns, int64_t T, int64_t n, void* stream) {
  if (n < 0 || T < 0 || !rewards || !values || !dones || !advantages || !returns) return set_err(BGYM_E_ARG, "bgym_gae: bad arguments");
  if (n == 0 || T == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  long long blocks = (n + 255) / 256;
  long long cap = (long long)g_sm_count * 32;
  gae_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(rewards, values, dones, gamma, lam, advantages, returns, T, n);
  return cuda_rc(cudaGetLastError(), "bgym_gae launch");
}

// ---- host-buffer handle API ---------------------------------------------------------------------
struct BgymVec {
  int64_t n; int device; cudaStream_t stream;
  uint8_t *d_hot, *d_tog, *d_cold, *d_obs, *d_sel, *d_term, *d_trunc, *d_decks; double* d_reward; BgymInfo* d_info; int32_t* d_actions;
  uint32_t* d_seeds; BgymDraws* d_draws;
  // pinned staging
  uint8_t *h_hot, *h_cold, *h_obs, *h_term, *h_trunc, *h_decks; double* h_reward; BgymInfo* h_info; int32_t* h_actions;
  uint32_t* h_seeds; BgymDraws* h_draws;
  // the step's outputs (obs | reward | info | terminated | truncated) live in ONE device block with a pinned mirror,
  // so a step is one device->host copy however many outputs are asked for
  uint8_t *d_block, *h_block; size_t block_bytes;
  uint8_t *d_mask, *h_mask;   // reset mask of bgym_vec_reset_masked_host
};

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return cuda_rc(_e, #x); } while (0)

// the handle entry points run on the handle's device and leave the caller's current device as they found it
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) { err = cudaGetDevice(&prev); if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define BGYM_ON_DEVICE(v) DeviceGuard _guard((v)->device); if (_guard.err != cudaSuccess) return cuda_rc(_guard.err, "cudaSetDevice")

static int vec_create_impl(BgymVec* v, int64_t n) {
  CK(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&v->d_hot, n * BGYM_HOT_BYTES)); CK(cudaMalloc(&v->d_cold, n * BGYM_COLD_BYTES));
  CK(cudaMalloc(&v->d_tog, n * BGYM_TOG_BYTES)); CK(cudaMalloc(&v->d_sel, n * BGYM_SEL_BYTES));
  CK(cudaMemsetAsync(v->d_tog, 0, n * BGYM_TOG_BYTES, v->stream)); CK(cudaMemsetAsync(v->d_sel, 0, n * BGYM_SEL_BYTES, v->stream));
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t off_rew = up16((size_t)n * BGYM_OBS_BYTES), off_info = up16(off_rew + (size_t)n * 8);
  const size_t off_term = up16(off_info + (size_t)n * BGYM_INFO_BYTES), off_trunc = up16(off_term + (size_t)n);
  v->block_bytes = up16(off_trunc + (size_t)n);
  CK(cudaMalloc(&v->d_block, v->block_bytes)); CK(cudaMallocHost(&v->h_block, v->block_bytes));
  v->d_obs = v->d_block; v->d_reward = reinterpret_cast<double*>(v->d_block + off_rew);
  v->d_info = reinterpret_cast<BgymInfo*>(v->d_block + off_info); v->d_term = v->d_block + off_term; v->d_trunc = v->d_block + off_trunc;
  v->h_obs = v->h_block; v->h_reward = reinterpret_cast<double*>(v->h_block + off_rew);
  v->h_info = reinterpret_cast<BgymInfo*>(v->h_block + off_info); v->h_term = v->h_block + off_term; v->h_trunc = v->h_block + off_trunc;
  CK(cudaMalloc(&v->d_decks, n * 52)); CK(cudaMalloc(&v->d_actions, n * 4));
  CK(cudaMalloc(&v->d_mask, n)); CK(cudaMallocHost(&v->h_mask, n));
  CK(cudaMalloc(&v->d_seeds, n * 4)); CK(cudaMalloc(&v->d_draws, n * BGYM_DRAWS_BYTES));
  CK(cudaMallocHost(&v->h_hot, n * BGYM_HOT_BYTES)); CK(cudaMallocHost(&v->h_cold, n * BGYM_COLD_BYTES));
  CK(cudaMallocHost(&v->h_decks, n * 52));
  CK(cudaMallocHost(&v->h_actions, n * 4)); CK(cudaMallocHost(&v->h_seeds, n * 4)); CK(cudaMallocHost(&v->h_draws, n * BGYM_DRAWS_BYTES));
  CK(cudaMemsetAsync(v->d_hot, 0, n * BGYM_HOT_BYTES, v->stream));
  CK(cudaMemsetAsync(v->d_cold, 0, n * BGYM_COLD_BYTES, v->stream));
  return 0;
}

int bgym_vec_create(BgymVec** out, int64_t n, int device) {
  if (!out || n <= 0) return set_err(BGYM_E_ARG, "bgym_vec_create: bad arguments");
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) return set_err(BGYM_E_NODEV, "bgym_vec_create: no CUDA device");
  if (device < 0 || device >= cnt) return set_err(BGYM_E_ARG, "bgym_vec_create: bad device index");
  BgymVec* v = new BgymVec();
  memset(v, 0, sizeof *v);
  v->n = n; v->device = device;
  int rc;
  {
    DeviceGuard guard(device);
    rc = guard.err != cudaSuccess ? cuda_rc(guard.err, "cudaSetDevice") : vec_create_impl(v, n);
  }
  if (rc) {                 // a failed allocation half way: give everything back (destroy frees what is there)
    char keep[sizeof g_err];
    memcpy(keep, g_err, sizeof keep);
    bgym_vec_destroy(v);
    memcpy(g_err, keep, sizeof keep);
    return rc;
  }
  *out = v;
  return 0;
}

int bgym_vec_destroy(BgymVec* v) {
  if (!v) return 0;
  DeviceGuard guard(v->device);
  if (v->stream) cudaStreamSynchronize(v->stream);
  cudaFree(v->d_hot); cudaFree(v->d_tog); cudaFree(v->d_cold); cudaFree(v->d_sel); cudaFree(v->d_block); cudaFree(v->d_decks); cudaFree(v->d_mask); cudaFreeHost(v->h_mask);
  cudaFree(v->d_actions); cudaFree(v->d_seeds); cudaFree(v->d_draws);
  cudaFreeHost(v->h_hot); cudaFreeHost(v->h_cold); cudaFreeHost(v->h_block);
  cudaFreeHost(v->h_decks); cudaFreeHost(v->h_actions);
  cudaFreeHost(v->h_seeds); cudaFreeHost(v->h_draws);
  if (v->stream) { bgym_release_stream(v->stream); cudaStreamDestroy(v->stream); }
  delete v;
  return 0;
}

int bgym_vec_reset_host(BgymVec* v, const uint32_t* seeds, const uint8_t* decks52, BgymObs* obs_out) {
  return bgym_vec_reset_masked_host(v, nullptr, seeds, decks52, obs_out);
}

int bgym_vec_reset_masked_host(BgymVec* v, const uint8_t* reset_mask, const uint32_t* seeds, const uint8_t* decks52, BgymObs* obs_out) {
  if (!v || !seeds) return set_err(BGYM_E_ARG, "bgym_vec_reset_host: bad arguments");
  BGYM_ON_DEVICE(v);
  if (reset_mask) {
    memcpy(v->h_mask, reset_mask, v->n);
    CK(cudaMemcpyAsync(v->d_mask, v->h_mask, v->n, cudaMemcpyHostToDevice, v->stream));
  }
  memcpy(v->h_seeds, seeds, v->n * 4);
  CK(cudaMemcpyAsync(v->d_seeds, v->h_seeds, v->n * 4, cudaMemcpyHostToDevice, v->stream));
  if (decks52) {
    memcpy(v->h_decks, decks52, v->n * 52);
    CK(cudaMemcpyAsync(v->d_decks, v->h_decks, v->n * 52, cudaMemcpyHostToDevice, v->stream));
  }
  int rc = bgym_reset(reinterpret_cast<BgymHot*>(v->d_hot), reinterpret_cast<BgymTog*>(v->d_tog), reinterpret_cast<BgymCold*>(v->d_cold),
                      reinterpret_cast<BgymObs*>(v->d_obs), reinterpret_cast<BgymSel*>(v->d_sel), reset_mask ? v->d_mask : nullptr, v->d_seeds,
                      decks52 ? v->d_decks : nullptr, v->n, 0, v->stream);
  if (rc) return rc;
  // a reset (masked or not) re-emits every env's observation record whole: no sync needed
  if (obs_out) CK(cudaMemcpyAsync(v->h_obs, v->d_obs, v->n * BGYM_OBS_BYTES, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  if (obs_out) memcpy(obs_out, v->h_obs, v->n * BGYM_OBS_BYTES);
  return 0;
}

int bgym_vec_step_host(BgymVec* v, const int32_t* actions, const BgymDraws* draws, BgymObs* obs_out,
                       double* reward_out, uint8_t* terminated_out, uint8_t* truncated_out,
                       BgymInfo* info_out, int flags) {
  if (!v || !actions) return set_err(BGYM_E_ARG, "bgym_vec_step_host: bad arguments");
  if (flags & BGYM_FLAG_RANDOM_POLICY)
    return set_err(BGYM_E_ARG, "bgym_vec_step_host: BGYM_FLAG_RANDOM_POLICY needs the device entry point (the host `actions` array is input only)");
  BGYM_ON_DEVICE(v);
  memcpy(v->h_actions, actions, v->n * 4);
  CK(cudaMemcpyAsync(v->d_actions, v->h_actions, v->n * 4, cudaMemcpyHostToDevice, v->stream));
  if (draws) {
    memcpy(v->h_draws, draws, v->n * BGYM_DRAWS_BYTES);
    CK(cudaMemcpyAsync(v->d_draws, v->h_draws, v->n * BGYM_DRAWS_BYTES, cudaMemcpyHostToDevice, v->stream));
  }
  int rc = bgym_step(reinterpret_cast<BgymHot*>(v->d_hot), reinterpret_cast<BgymTog*>(v->d_tog), reinterpret_cast<BgymCold*>(v->d_cold), v->d_actions,
                     draws ? v->d_draws : nullptr, reinterpret_cast<BgymObs*>(v->d_obs), reinterpret_cast<BgymSel*>(v->d_sel), nullptr,
                     v->d_reward, v->d_term, v->d_trunc, v->d_info, v->n, flags & ~BGYM_FLAG_NO_OBS, v->stream);
  if (rc) return rc;
  if (obs_out) {
    // host callers get whole records: fold the selection array (selected_cards, mask word) into them first — unless the slab
    // took the one-launch step, which rewrites every record whole (the N = 1 Gymnasium facade: one launch less per step)
    static const bool diag = getenv("BGYM_STEP_TIMING") != nullptr;
    if (v->n > small_slab_threshold() || diag) {
      rc = bgym_sync_obs(reinterpret_cast<BgymObs*>(v->d_obs), reinterpret_cast<BgymSel*>(v->d_sel), v->n, BGYM_SYNC_TO_RECORDS, v->stream);
      if (rc) return rc;
    }
    CK(cudaMemcpyAsync(v->h_block, v->d_block, v->block_bytes, cudaMemcpyDeviceToHost, v->stream));      // everything, one copy
  } else {   // without observations: skip the 176 B/env part
    const size_t off = reinterpret_cast<uint8_t*>(v->d_reward) - v->d_block;
    CK(cudaMemcpyAsync(v->h_block + off, v->d_block + off, v->block_bytes - off, cudaMemcpyDeviceToHost, v->stream));
  }
  CK(cudaStreamSynchronize(v->stream));
  if (obs_out) memcpy(obs_out, v->h_obs, v->n * BGYM_OBS_BYTES);
  if (reward_out) memcpy(reward_out, v->h_reward, v->n * 8);
  if (terminated_out) memcpy(terminated_out, v->h_term, v->n);
  if (truncated_out) memcpy(truncated_out, v->h_trunc, v->n);
  if (info_out) memcpy(info_out, v->h_info, v->n * BGYM_INFO_BYTES);
  return 0;
}

int bgym_vec_pointers(BgymVec* v, void** hot, void** tog, void** cold, void** obs, void** sel, void** reward, void** terminated) {
  if (!v) return set_err(BGYM_E_ARG, "bgym_vec_pointers: null handle");
  if (hot) *hot = v->d_hot;
  if (tog) *tog = v->d_tog;
  if (cold) *cold = v->d_cold;
  if (obs) *obs = v->d_obs;
  if (sel) *sel = v->d_sel;
  if (reward) *reward = v->d_reward;
  if (terminated) *terminated = v->d_term;
  return 0;
}

// host-side BgymState records = {hot, cold} back to back
int bgym_vec_get_state(BgymVec* v, BgymState* host_out) {
  if (!v || !host_out) return set_err(BGYM_E_ARG, "bgym_vec_get_state: bad arguments");
  BGYM_ON_DEVICE(v);
  int rc = bgym_sync_state(reinterpret_cast<BgymHot*>(v->d_hot), reinterpret_cast<BgymTog*>(v->d_tog), v->n, BGYM_SYNC_TO_RECORDS, v->stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(v->h_hot, v->d_hot, v->n * BGYM_HOT_BYTES, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaMemcpyAsync(v->h_cold, v->d_cold, v->n * BGYM_COLD_BYTES, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  uint8_t* o = reinterpret_cast<uint8_t*>(host_out);
  for (int64_t i = 0; i < v->n; i++) {
    memcpy(o + i * BGYM_STATE_BYTES, v->h_hot + i * BGYM_HOT_BYTES, BGYM_HOT_BYTES);
    memcpy(o + i * BGYM_STATE_BYTES + BGYM_HOT_BYTES, v->h_cold + i * BGYM_COLD_BYTES, BGYM_COLD_BYTES);
  }
  return 0;
}

int bgym_vec_set_state(BgymVec* v, const BgymState* host_in) {
  if (!v || !host_in) return set_err(BGYM_E_ARG, "bgym_vec_set_state: bad arguments");
  BGYM_ON_DEVICE(v);
  const uint8_t* in = reinterpret_cast<const uint8_t*>(host_in);
  for (int64_t i = 0; i < v->n; i++) {
    memcpy(v->h_hot + i * BGYM_HOT_BYTES, in + i * BGYM_STATE_BYTES, BGYM_HOT_BYTES);
    memcpy(v->h_cold + i * BGYM_COLD_BYTES, in + i * BGYM_STATE_BYTES + BGYM_HOT_BYTES, BGYM_COLD_BYTES);
  }
  CK(cudaMemcpyAsync(v->d_hot, v->h_hot, v->n * BGYM_HOT_BYTES, cudaMemcpyHostToDevice, v->stream));
  CK(cudaMemcpyAsync(v->d_cold, v->h_cold, v->n * BGYM_COLD_BYTES, cudaMemcpyHostToDevice, v->stream));
  int rc = bgym_sync_state(reinterpret_cast<BgymHot*>(v->d_hot), reinterpret_cast<BgymTog*>(v->d_tog), v->n, BGYM_SYNC_FROM_RECORDS, v->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(v->stream));
  return 0;
}

}  // extern "C"
