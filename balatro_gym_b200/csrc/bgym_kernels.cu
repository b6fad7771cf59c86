// bgym_kernels.cu — sm_100a kernels and the C-ABI (include/bgym.h) of the batched Balatro env.
//
// Kernels
//   env_step_kernel     K2  fused BalatroEnv.step: mask check -> phase dispatch -> scoring -> boss ->
//                           round advance / shop generation -> reward -> observation + mask emission,
//                           optional in-place autoreset (K3) and fused random-legal policy.
//   env_reset_kernel    K3  reset + first observation (same staging as K2)
//   score_hands_kernel  K1  classify + chips x mult + joker interpreter (K4) per hand
//   action_mask_kernel, sample_actions_kernel, episode_stats_kernel (K6)
//
// Data movement of K2/K3: state is a dense array of 320 B records.  Each warp owns a tile of 32
// envs; every lane pulls its own record into shared memory with ONE 1-D bulk async copy
// (cp.async.bulk, SASS UBLKCP — the TMA engine without a tensor map) that completes on the warp's
// mbarrier, works on it in place, assembles the 240 B observation record in shared memory with
// 128-bit stores, and pushes both back with bulk async stores.  Shared-memory strides are odd
// multiples of 16 B (336 / 240) so the 128-bit accesses of a quarter-warp hit distinct bank groups.
// Tiles are double buffered per warp: the bulk loads of the next tile are in flight while the
// current one is computed.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bgym_env.cuh"

namespace bgym {

constexpr int REC_STRIDE = BGYM_STATE_BYTES;  // 304 = 19 x 16 B: odd 16-byte stride, no padding needed
constexpr int OBS_STRIDE = BGYM_OBS_BYTES;    // 240 = 15 x 16 B
constexpr int WARP_STATE_BYTES = 32 * REC_STRIDE;              // 9728
constexpr int WARP_OBS_BYTES = 32 * OBS_STRIDE;                // 7680

// Two staging variants (picked at run time, BGYM_VARIANT=0/1):
//   V0: single state buffer per warp, 4 warps/CTA, 3 CTAs/SM  -> 12 warps/SM hide each other's loads
//   V1: double-buffered state per warp, 8 warps/CTA, 1 CTA/SM -> 8 warps/SM, next tile prefetched
//   V6/V7: single state buffer, observation written straight to global with 128-bit stores (no obs
//          staging) -> 9.7 KB of shared memory per warp, 20+ warps/SM
template <int STAGES, int WARPS, bool OBS_DIRECT = false, int CTAS = 0>
struct Cfg {
  static constexpr int stages = STAGES, warps = WARPS, threads = WARPS * 32;
  static constexpr bool obs_direct = OBS_DIRECT;
  static constexpr int warp_smem = STAGES * WARP_STATE_BYTES + (OBS_DIRECT ? 0 : WARP_OBS_BYTES);
  static constexpr int cta_smem = WARPS * warp_smem + 16 * WARPS;  // + mbarriers
  static constexpr int ctas_per_sm = CTAS ? CTAS : (227 * 1024) / cta_smem;
};
using CfgV0 = Cfg<1, 4>;
using CfgV1 = Cfg<2, 8>;
using CfgV6 = Cfg<1, 4, true, 5>;
using CfgV7 = Cfg<1, 3, true, 7>;

struct StepArgs {
  uint8_t* state;            // n x 320
  const int32_t* actions;    // n (nullable with BGYM_FLAG_RANDOM_POLICY)
  int32_t* actions_out;      // n (written with BGYM_FLAG_RANDOM_POLICY, nullable)
  const BgymDraws* draws;    // n (nullable)
  uint8_t* obs;              // n x 240 (nullable)
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
  // partitioned step: three device lists of deferred env indices + their counters
  int* part_lists;           // [3][part_cap]
  int* part_counters;        // [4] (index c = category c; 0 unused)
  long long part_cap;
};

enum { MODE_STEP = 0, MODE_RESET = 1 };

}  // namespace bgym
#include "bgym_step_sorted.cuh"
#include "bgym_step_part.cuh"
namespace bgym {

template <int MODE, typename C>
__global__ void __launch_bounds__(C::threads, C::ctas_per_sm) env_kernel(StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int STAGES = C::stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wbase = smem + warp * C::warp_smem;
  uint8_t* obs_buf = wbase + STAGES * WARP_STATE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::warps * C::warp_smem) + warp * 2;

  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncwarp();

  const long long n_tiles = (a.n + 31) >> 5;
  const long long warp_gid = (long long)blockIdx.x * C::warps + warp;
  const long long warp_cnt = (long long)gridDim.x * C::warps;
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);

  // a reset of ALL envs needs no state load at all
  const bool need_load = (MODE == MODE_STEP) || (a.reset_mask != nullptr);

  auto issue_load = [&](long long tile, int stage) {
    // ONE bulk copy per warp tile: 32 consecutive records are contiguous in global and in shared
    if (lane == 0) {
      uint32_t bytes = (uint32_t)(min(32LL, a.n - tile * 32) * BGYM_STATE_BYTES);
      mbar_arrive_expect_tx(&bars[stage], bytes);
      bulk_g2s(wbase + stage * WARP_STATE_BYTES, a.state + tile * 32 * BGYM_STATE_BYTES, bytes, &bars[stage]);
    }
  };

  long long tile = warp_gid;
  int stage = 0;
  uint32_t phase_bits = 0;  // parity per stage
  if (STAGES == 2 && need_load && tile < n_tiles) issue_load(tile, 0);

  for (; tile < n_tiles; tile += warp_cnt, stage = (STAGES == 2) ? (stage ^ 1) : 0) {
    const long long e = tile * 32 + lane;
    const bool active = e < a.n;
    uint8_t* rec = wbase + stage * WARP_STATE_BYTES + lane * REC_STRIDE;
    uint8_t* obs_s = C::obs_direct ? (a.obs + e * BGYM_OBS_BYTES) : (obs_buf + lane * OBS_STRIDE);

    // the buffer about to be overwritten (and the obs buffer) were last READ by the bulk stores
    // lane 0 issued in the previous iteration: wait for those reads to finish
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (STAGES == 2) {
      const long long next = tile + warp_cnt;
      if (need_load && next < n_tiles) issue_load(next, stage ^ 1);
    } else {
      if (need_load) issue_load(tile, 0);
    }

    // inputs that do not depend on the state: fetch while the bulk copy lands
    int action = 0;
    if (MODE == MODE_STEP && active && a.actions && !(a.flags & BGYM_FLAG_RANDOM_POLICY)) action = __ldg(a.actions + e);

    if (need_load) {
      mbar_wait(&bars[stage], (phase_bits >> stage) & 1);
      phase_bits ^= 1u << stage;
    }

    Hot h;
    double reward = 0.0;
    int terminated = 0;
    StepInfo info;
    bool do_store_state = true, want_reset = false;
    uint32_t new_seed = 0;
    if (active) {
      if (MODE == MODE_STEP) {
        unpack_hot(rec, h);
        uint64_t m0 = action_mask(h, rec);
        if (a.flags & BGYM_FLAG_RANDOM_POLICY) {
          // uniform legal action: Philox keyed by (seed, policy key), counter = episode step
          int cnt = __popcll(m0);
          uint4 w = philox4x32_10(h.ep_len, 0, 0, 0, h.rng_seed, BGYM_POLICY_KEY1);
          int k = (int)__umulhi(w.x, (uint32_t)cnt);
          uint64_t mm = m0;
          for (int i = 0; i < k; i++) mm &= mm - 1;
          action = cnt ? __ffsll((long long)mm) - 1 : 0;
          if (a.actions_out) a.actions_out[e] = action;
        }
        step_env<CAT_ALL>(h, rec, action, m0, a.draws ? a.draws + e : nullptr, reward, terminated, info);
        if (terminated && (a.flags & BGYM_FLAG_AUTORESET)) {
          uint32_t episode = h.episode + 1;
          new_seed = next_episode_seed(h.rng_seed);
          reset_hot(h, new_seed);
          h.episode = episode;
          info.flags |= BGYM_F_AUTORESET_DONE;
          want_reset = true;
        }
      } else {
        if (a.reset_mask && !a.reset_mask[e]) {
          unpack_hot(rec, h);   // untouched env: only re-emit its observation
          do_store_state = false;
        } else {
          reset_hot(h, a.seeds[e]);
          reset_blocks_serial(rec, a.seeds[e], a.decks52 ? a.decks52 + e * 52 : nullptr);
        }
      }
    }
    if (MODE == MODE_STEP && (a.flags & BGYM_FLAG_AUTORESET)) {
      // in-place autoreset: the deck/shop blocks of every terminated env of the tile are rebuilt by
      // the WHOLE warp, one env after the other (few lanes terminate per step)
      autoreset_warp(want_reset, new_seed, rec, lane);
    }
    if (active) {
      if (do_store_state) pack_hot(rec, h);
      if (with_obs) write_obs(h, rec, action_mask(h, rec), obs_s);
      fence_async_smem();  // this lane's shared-memory writes -> visible to the async proxy
      if (MODE == MODE_STEP) {
        a.reward[e] = reward;
        a.terminated[e] = (uint8_t)terminated;
        if (a.truncated) a.truncated[e] = 0;
        if (a.info) {
          uint4 i0, i1;
          i0.x = (uint32_t)info.final_score; i0.y = (uint32_t)((uint64_t)info.final_score >> 32);
          unsigned long long xb = (unsigned long long)__double_as_longlong(info.x_mult);
          i0.z = (uint32_t)xb; i0.w = (uint32_t)(xb >> 32);
          i1.x = (uint32_t)info.chips; i1.y = (uint32_t)info.mult;
          i1.z = (uint32_t)(info.hand_type & 0xFF) | ((uint32_t)(info.error_code & 0xFF) << 8) |
                 ((uint32_t)(info.flags & 0xFF) << 16) | ((uint32_t)(info.cards_played & 0xFF) << 24);
          i1.w = (uint32_t)info.base_score;
          uint4* ip = reinterpret_cast<uint4*>(a.info + e);
          ip[0] = i0; ip[1] = i1;
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      // ONE bulk store per warp tile for the state records and one for the observation records
      uint32_t cnt = (uint32_t)min(32LL, a.n - tile * 32);
      bulk_s2g(a.state + tile * 32 * BGYM_STATE_BYTES, wbase + stage * WARP_STATE_BYTES, cnt * BGYM_STATE_BYTES);
      if (with_obs && !C::obs_direct) bulk_s2g(a.obs + tile * 32 * BGYM_OBS_BYTES, obs_buf, cnt * BGYM_OBS_BYTES);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait0();  // every bulk store of this warp has completed before exit
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
  uint32_t seed, ctr; unsigned long long index; uint4 buf; int pos;
  __device__ __forceinline__ uint32_t word() {
    if (pos == 4) { buf = philox4x32_10(ctr++, (uint32_t)index, (uint32_t)(index >> 32), 1, seed, BGYM_PHILOX_KEY1); pos = 0; }
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

// Fast path of config 2 (five-card plays, no modifiers, no jokers, level-1 hands): everything is
// static — five byte extracts, register histograms, one table lookup — and each thread scores two
// hands per iteration so two independent load->compute->store chains are in flight.
__global__ void __launch_bounds__(256) score_hands5_kernel(ScoreArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    uint2 cw = __ldg(reinterpret_cast<const uint2*>(a.cards8) + i);
    int c0 = cw.x & 0xFF, c1 = (cw.x >> 8) & 0xFF, c2 = (cw.x >> 16) & 0xFF, c3 = cw.x >> 24, c4 = cw.y & 0xFF;
    HandHist hist;
    hist.clear();
    hist.add(c0); hist.add(c1); hist.add(c2); hist.add(c3); hist.add(c4);
    int chip_sum = card_chips(c0, 0, 0) + card_chips(c1, 0, 0) + card_chips(c2, 0, 0) + card_chips(c3, 0, 0) + card_chips(c4, 0, 0);
    int ht = classify(hist);
    int chips = c_base_chips[ht] + chip_sum, mult = c_base_mult[ht];
    a.hand_type[i] = (uint8_t)ht;
    a.chips[i] = chips;
    a.mult[i] = mult;
    a.score[i] = (long long)chips * mult;   // x_mult == 1.0: int(chips * mult * 1.0)
    if (a.x_mult) a.x_mult[i] = 1.0;
    if (a.money) a.money[i] = 0;
  }
}

__global__ void __launch_bounds__(256) score_hands_kernel(ScoreArgs a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    uint2 cw = __ldg(reinterpret_cast<const uint2*>(a.cards8) + i);
    uint64_t cards = u64_of(cw.x, cw.y);
    int nc = a.n_cards ? a.n_cards[i] : 5;
    uint4 mw = make_uint4(0, 0, 0, 0);
    if (a.mods8) mw = __ldg(reinterpret_cast<const uint4*>(a.mods8) + i);
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
    int ht = classify(hist);
    int lvl = a.levels12 ? a.levels12[i * 12 + ht] : 1;
    int chips, mult;
    hand_base(ht, lvl, chips, mult);
    chips += chip_sum;
    double x_mult = 1.0;
    int money = 0;
    if (a.jokers8) {
      // K4: table-driven joker interpreter (unified_scoring.py:156-244), joker order preserved
      uint2 jw = __ldg(reinterpret_cast<const uint2*>(a.jokers8) + i);
      uint64_t jraw = u64_of(jw.x, jw.y), jk = 0;
      int nj = 0;
      for (int j = 0; j < 8; j++) { int id = byte_at(jraw, j); if (id) { jk |= (uint64_t)id << (8 * nj); nj++; } }
      BgymScoreCtx cx;
      cx.hands_left = 4; cx.discards_left = 3; cx.deck_len = 52; cx.use_replay = 0; cx.bloodstone_bits = 0;
      cx.misprint[0] = 0;
      if (a.ctx) cx = a.ctx[i];
      bool table_names = (a.flags & BGYM_SCORE_TABLE_NAMES) != 0;
      ScoreRng rng; rng.seed = a.seed; rng.ctr = 0; rng.index = (unsigned long long)i; rng.pos = 4;
      // individual phase: card-major, joker-minor (:173-209)
      int ind_chips = 0, ind_mult = 0; double ind_x = 1.0;
      for (int c = 0; c < nc; c++) {
        int rank = byte_at(rank8, c), suit = nib_at(suit8, c);
        for (int j = 0; j < nj; j++) {
          const BgymJokerFx fx = c_joker_fx[byte_at(jk, j)];
          bool fire = false;
          if (fx.kind == BGYM_FX_IND_RANKSET || fx.kind == BGYM_FX_IND_FACE) fire = (fx.arg >> rank) & 1;
          else if (fx.kind == BGYM_FX_IND_SUIT) {
            fire = suit == (fx.arg & 3);
            if (fx.arg & 0x80) {  // Bloodstone: one roll per (card, joker) pair (:161)
              bool hit = cx.use_replay ? ((cx.bloodstone_bits >> c) & 1) : (rng.u01() < 0.5);
              fire = fire && hit;
            }
          }
          if (fire) { ind_chips += fx.chips; ind_mult += fx.mult; ind_x *= (double)fx.xmult; money += fx.money; }
        }
      }
      chips += ind_chips; mult += ind_mult; x_mult *= ind_x;
      // main phase, joker order (:211-244)
      int misprint_seen = 0;
      int n_suit_names = __popc(suit_present) + (int)stone_present;
      for (int j = 0; j < nj; j++) {
        const BgymJokerFx fx = c_joker_fx[byte_at(jk, j)];
        int ec = 0, em = 0; double ex = 1.0;
        switch (fx.kind) {
          case BGYM_FX_MAIN_ALWAYS: ec = fx.chips; em = fx.mult; ex = fx.xmult; break;
          case BGYM_FX_MAIN_HANDNAME: if (name_matches(ht, table_names, fx.arg)) { ec = fx.chips; em = fx.mult; ex = fx.xmult; } break;
          case BGYM_FX_MAIN_SUIT_ANY: if ((suit_present >> fx.arg) & 1) { em = fx.mult; } break;
          case BGYM_FX_MAIN_STATE:
            switch (fx.arg) {
              case BGYM_FXS_HALF: if (nc <= 3) em = fx.mult; break;
              case BGYM_FXS_ABSTRACT: em = fx.mult * nj; break;
              case BGYM_FXS_ACROBAT: if (cx.hands_left == 1) ex = fx.xmult; break;
              case BGYM_FXS_MYSTIC: if (cx.discards_left == 0) em = fx.mult; break;
              case BGYM_FXS_BANNER: ec = fx.chips * cx.discards_left; break;
              case BGYM_FXS_BLUE: ec = fx.chips * cx.deck_len; break;
              case BGYM_FXS_MISPRINT:
                em = cx.use_replay ? cx.misprint[min(misprint_seen, 4)] : rng.below(24);
                misprint_seen++;
                break;
            }
            break;
          case BGYM_FX_MAIN_SPECIAL:
            switch (fx.arg) {
              case BGYM_FXSP_BLACKBOARD: if (all_black) ex = fx.xmult; break;
              case BGYM_FXSP_SEEING_DOUBLE: if ((suit_present & 1) && n_suit_names > 1) ex = fx.xmult; break;
              case BGYM_FXSP_FLOWER_POT: if (n_suit_names == 4) ex = fx.xmult; break;
              case BGYM_FXSP_BARON: if (kings > 0) ex = c_pow_1_5[kings]; break;
              case BGYM_FXSP_SHOOT_MOON: if (queens > 0) em = fx.mult * queens; break;
            }
            break;
          default: break;
        }
        chips += ec; mult += em; x_mult *= ex;
      }
    }
    // final_score = int(chips * mult * x_mult), unified_scoring.py:286
    long long score = (long long)((double)((long long)chips * (long long)mult) * x_mult);
    a.hand_type[i] = (uint8_t)ht;
    a.chips[i] = chips;
    a.mult[i] = mult;
    if (a.x_mult) a.x_mult[i] = x_mult;
    a.score[i] = score;
    if (a.money) a.money[i] = money;
  }
}

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void action_mask_kernel(const uint8_t* state, uint64_t* mask, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* rec = state + i * BGYM_STATE_BYTES;
    uint64_t m = 0;
    int phase = rec[17];
    if (phase == BGYM_PHASE_PLAY) {
      int hand_n = rec[8], sel_n = rec[10], discards_left = rec[21], cons_n = rec[23];
      m = ((1ull << min(hand_n, 8)) - 1) << BGYM_A_SELECT_BASE;
      if (sel_n > 0) m |= 1ull << BGYM_A_PLAY_HAND;
      if (sel_n > 0 && discards_left > 0) m |= 1ull << BGYM_A_DISCARD;
      m |= ((1ull << cons_n) - 1) << BGYM_A_USE_CONS_BASE;
    } else if (phase == BGYM_PHASE_SHOP) {
      int money = *reinterpret_cast<const int*>(rec + 32), joker_n = rec[22];
      int n_items = rec[OFF_N_ITEMS];
      for (int k = 0; k < n_items; k++)
        if (money >= *reinterpret_cast<const int*>(rec + OFF_ITEM_COST + 4 * k)) m |= 1ull << (BGYM_A_SHOP_BUY_BASE + k);
      if (money >= *reinterpret_cast<const int*>(rec + 108)) m |= 1ull << BGYM_A_SHOP_REROLL;
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
__global__ void sample_actions_kernel(const uint8_t* obs, int32_t* actions, uint32_t seed, unsigned long long step, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint64_t m = *reinterpret_cast<const uint64_t*>(obs + i * BGYM_OBS_BYTES + 160);
    int cnt = __popcll(m);
    int act = 0;
    if (cnt) {
      uint4 w = philox4x32_10((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), (uint32_t)step, (uint32_t)(step >> 32), seed, BGYM_POLICY_KEY1);
      int k = (int)__umulhi(w.x, (uint32_t)cnt);
      for (int t = 0; t < k; t++) m &= m - 1;
      act = __ffsll((long long)m) - 1;
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
  // warp reduce, one atomic per warp and statistic
  for (int o = 16; o > 0; o >>= 1) {
    eps += __shfl_xor_sync(0xffffffffu, eps, o); sret += __shfl_xor_sync(0xffffffffu, sret, o);
    slen += __shfl_xor_sync(0xffffffffu, slen, o); steps += __shfl_xor_sync(0xffffffffu, steps, o);
    srew += __shfl_xor_sync(0xffffffffu, srew, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (eps != 0) { atomicAdd(&stats[0], eps); atomicAdd(&stats[1], sret); atomicAdd(&stats[2], slen); }
    atomicAdd(&stats[3], steps); atomicAdd(&stats[4], srew);
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

static int g_sm_count = 0;
static bool g_attr_set = false;
// step-kernel variants (BGYM_VARIANT): 0/1 fixed lane<->env mapping (single / double buffered),
// 2..5 CTA-level path sorting with 4x3, 6x2, 8x1, 12x1 (warps per CTA x CTAs per SM)
//   8 category-partitioned step (main pass + one gather pass per rare category)  <- default
#define BGYM_DEFAULT_VARIANT 8
using SortedA = SortedCfg<4, 3>;
using SortedB = SortedCfg<6, 2>;
using SortedC = SortedCfg<8, 1>;
using SortedD = SortedCfg<12, 1>;
static int g_variant = BGYM_DEFAULT_VARIANT;
static int ensure_device_setup() {
  if (g_attr_set) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_rc(e, "cudaGetDevice");
  e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_rc(e, "cudaDeviceGetAttribute");
  e = cudaFuncSetAttribute(env_kernel<MODE_STEP, CfgV0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgV0::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(step v0)");
  e = cudaFuncSetAttribute(env_kernel<MODE_RESET, CfgV0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgV0::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(reset v0)");
  e = cudaFuncSetAttribute(env_kernel<MODE_STEP, CfgV1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgV1::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(step v1)");
  e = cudaFuncSetAttribute(env_kernel<MODE_STEP, CfgV6>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgV6::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(step v6)");
  e = cudaFuncSetAttribute(env_kernel<MODE_STEP, CfgV7>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgV7::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(step v7)");
  e = cudaFuncSetAttribute(env_step_sorted_kernel<SortedA>, cudaFuncAttributeMaxDynamicSharedMemorySize, SortedA::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(sorted A)");
  e = cudaFuncSetAttribute(env_step_sorted_kernel<SortedB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SortedB::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(sorted B)");
  e = cudaFuncSetAttribute(env_step_sorted_kernel<SortedC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SortedC::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(sorted C)");
  e = cudaFuncSetAttribute(env_step_sorted_kernel<SortedD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SortedD::cta_smem);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(sorted D)");
  e = cudaFuncSetAttribute(env_step_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PART_CTA_SMEM);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(part main)");
  e = cudaFuncSetAttribute(env_step_gather_kernel<CAT_PLAY, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, PART_CTA_SMEM);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(part play)");
  e = cudaFuncSetAttribute(env_step_gather_kernel<CAT_DISCARD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PART_CTA_SMEM);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(part discard)");
  e = cudaFuncSetAttribute(env_step_gather_kernel<CAT_OTHER, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PART_CTA_SMEM);
  if (e != cudaSuccess) return cuda_rc(e, "cudaFuncSetAttribute(part other)");
  const char* v = getenv("BGYM_VARIANT");
  g_variant = (v && v[0] >= '0' && v[0] <= '8') ? (v[0] - '0') : BGYM_DEFAULT_VARIANT;
  g_attr_set = true;
  return 0;
}

// scratch of the partitioned step (deferred-env lists + counters), one per (device, stream)
struct PartScratch { int dev; void* stream; long long cap; int* lists; int* counters;
                     cudaStream_t side[2]; cudaEvent_t ev_main, ev_side[2]; bool streams_ok; };
static PartScratch g_scratch[16];
static int g_n_scratch = 0;
static int get_part_scratch(long long n, void* stream, PartScratch** out) {
  int dev = 0;
  cudaGetDevice(&dev);
  PartScratch* sc = nullptr;
  for (int i = 0; i < g_n_scratch; i++)
    if (g_scratch[i].dev == dev && g_scratch[i].stream == stream) sc = &g_scratch[i];
  if (!sc) {
    if (g_n_scratch == 16) return set_err(BGYM_E_ARG, "bgym_step: too many (device, stream) pairs in use");
    sc = &g_scratch[g_n_scratch++];
    sc->dev = dev; sc->stream = stream; sc->cap = 0; sc->lists = nullptr; sc->counters = nullptr;
    // the three gather passes touch disjoint envs and are latency-bound: run them concurrently
    sc->streams_ok = cudaStreamCreateWithFlags(&sc->side[0], cudaStreamNonBlocking) == cudaSuccess &&
                     cudaStreamCreateWithFlags(&sc->side[1], cudaStreamNonBlocking) == cudaSuccess &&
                     cudaEventCreateWithFlags(&sc->ev_main, cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&sc->ev_side[0], cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&sc->ev_side[1], cudaEventDisableTiming) == cudaSuccess;
  }
  if (sc->cap < n) {
    if (sc->lists) cudaFree(sc->lists);
    cudaError_t e = cudaMalloc(&sc->lists, (size_t)(3 * n + 4) * sizeof(int));
    if (e != cudaSuccess) { sc->cap = 0; sc->lists = nullptr; return cuda_rc(e, "cudaMalloc(step scratch)"); }
    sc->cap = n;
    sc->counters = sc->lists + 3 * n;
  }
  *out = sc;
  return 0;
}

static int launch_partitioned(StepArgs& a, cudaStream_t s) {
  PartScratch* sc = nullptr;
  int rc = get_part_scratch(a.n, (void*)s, &sc);
  if (rc) return rc;
  a.part_lists = sc->lists; a.part_counters = sc->counters; a.part_cap = sc->cap;
  cudaError_t e = cudaMemsetAsync(sc->counters, 0, 4 * sizeof(int), s);
  if (e != cudaSuccess) return cuda_rc(e, "cudaMemsetAsync(step counters)");
  long long tiles = (a.n + 31) / 32;
  long long ctas = (tiles + PART_WARPS - 1) / PART_WARPS;
  long long cap = (long long)g_sm_count * PART_CTAS_PER_SM;
  int grid = (int)(ctas < cap ? ctas : cap);
  env_step_main_kernel<<<grid, PART_WARPS * 32, PART_CTA_SMEM, s>>>(a);
  // the list lengths live on the device: launch resident-size grids, idle warps exit at once
  long long gcap = (ctas + 3) / 4 < cap ? (ctas + 3) / 4 : cap;
  int ggrid = (int)(gcap < 1 ? 1 : gcap);
  static const bool serial = getenv("BGYM_SERIAL_GATHER") != nullptr;
  if (sc->streams_ok && !serial) {
    cudaEventRecord(sc->ev_main, s);
    cudaStreamWaitEvent(sc->side[0], sc->ev_main, 0);
    env_step_gather_kernel<CAT_PLAY, 0><<<ggrid, PART_WARPS * 32, PART_CTA_SMEM, sc->side[0]>>>(a);
    cudaEventRecord(sc->ev_side[0], sc->side[0]);
    cudaStreamWaitEvent(sc->side[1], sc->ev_main, 0);
    env_step_gather_kernel<CAT_OTHER, 2><<<ggrid, PART_WARPS * 32, PART_CTA_SMEM, sc->side[1]>>>(a);
    cudaEventRecord(sc->ev_side[1], sc->side[1]);
    env_step_gather_kernel<CAT_DISCARD, 1><<<ggrid, PART_WARPS * 32, PART_CTA_SMEM, s>>>(a);
    cudaStreamWaitEvent(s, sc->ev_side[0], 0);
    cudaStreamWaitEvent(s, sc->ev_side[1], 0);
  } else {
    env_step_gather_kernel<CAT_PLAY, 0><<<ggrid, PART_WARPS * 32, PART_CTA_SMEM, s>>>(a);
    env_step_gather_kernel<CAT_DISCARD, 1><<<ggrid, PART_WARPS * 32, PART_CTA_SMEM, s>>>(a);
    env_step_gather_kernel<CAT_OTHER, 2><<<ggrid, PART_WARPS * 32, PART_CTA_SMEM, s>>>(a);
  }
  return 0;
}

template <typename C>
static void launch_sorted(const StepArgs& a, cudaStream_t s) {
  long long tiles = (a.n + C::T - 1) / C::T;
  long long cap = (long long)g_sm_count * C::ctas_per_sm;
  env_step_sorted_kernel<C><<<(int)(tiles < cap ? tiles : cap), C::threads, C::cta_smem, s>>>(a);
}

template <typename C>
static int env_grid(long long n) {
  long long tiles = (n + 31) / 32;
  long long ctas = (tiles + C::warps - 1) / C::warps;
  long long cap = (long long)g_sm_count * C::ctas_per_sm;  // persistent: resident CTAs only
  return (int)(ctas < cap ? ctas : cap);
}

static bool misaligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

extern "C" {

int bgym_abi_version(void) { return BGYM_ABI_VERSION; }
const char* bgym_last_error(void) { return g_err; }
int bgym_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int bgym_reset(BgymState* state, BgymObs* obs, const uint8_t* reset_mask, const uint32_t* seeds,
               const uint8_t* decks52, int64_t n, int flags, void* stream) {
  if (n < 0 || !state || !seeds) return set_err(BGYM_E_ARG, "bgym_reset: null state/seeds or negative n");
  if (!obs && !(flags & BGYM_FLAG_NO_OBS)) return set_err(BGYM_E_ARG, "bgym_reset: obs is NULL without BGYM_FLAG_NO_OBS");
  if (misaligned(state, 16) || misaligned(obs, 16)) return set_err(BGYM_E_ALIGN, "bgym_reset: state/obs must be 16-byte aligned");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.state = reinterpret_cast<uint8_t*>(state); a.obs = reinterpret_cast<uint8_t*>(obs);
  a.reset_mask = reset_mask; a.seeds = seeds; a.decks52 = decks52; a.n = n; a.flags = flags;
  env_kernel<MODE_RESET, CfgV0><<<env_grid<CfgV0>(n), CfgV0::threads, CfgV0::cta_smem, (cudaStream_t)stream>>>(a);
  return cuda_rc(cudaGetLastError(), "bgym_reset launch");
}

int bgym_step(BgymState* state, int32_t* actions, const BgymDraws* draws, BgymObs* obs,
              double* reward, uint8_t* terminated, uint8_t* truncated, BgymInfo* info,
              int64_t n, int flags, void* stream) {
  if (n < 0 || !state || !reward || !terminated) return set_err(BGYM_E_ARG, "bgym_step: null state/reward/terminated or negative n");
  if (!actions) return set_err(BGYM_E_ARG, "bgym_step: actions is NULL");
  if (!obs && !(flags & BGYM_FLAG_NO_OBS)) return set_err(BGYM_E_ARG, "bgym_step: obs is NULL without BGYM_FLAG_NO_OBS");
  if (misaligned(state, 16) || misaligned(obs, 16) || misaligned(info, 16) || misaligned(draws, 8))
    return set_err(BGYM_E_ALIGN, "bgym_step: state/obs/info must be 16-byte aligned");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.state = reinterpret_cast<uint8_t*>(state);
  a.actions = actions;
  a.actions_out = (flags & BGYM_FLAG_RANDOM_POLICY) ? actions : nullptr;
  a.draws = draws; a.obs = reinterpret_cast<uint8_t*>(obs);
  a.reward = reward; a.terminated = terminated; a.truncated = truncated; a.info = info;
  a.n = n; a.flags = flags;
  cudaStream_t s = (cudaStream_t)stream;
  switch (g_variant) {
    case 0: env_kernel<MODE_STEP, CfgV0><<<env_grid<CfgV0>(n), CfgV0::threads, CfgV0::cta_smem, s>>>(a); break;
    case 1: env_kernel<MODE_STEP, CfgV1><<<env_grid<CfgV1>(n), CfgV1::threads, CfgV1::cta_smem, s>>>(a); break;
    case 2: launch_sorted<SortedA>(a, s); break;
    case 4: launch_sorted<SortedC>(a, s); break;
    case 5: launch_sorted<SortedD>(a, s); break;
    case 6: env_kernel<MODE_STEP, CfgV6><<<env_grid<CfgV6>(n), CfgV6::threads, CfgV6::cta_smem, s>>>(a); break;
    case 7: env_kernel<MODE_STEP, CfgV7><<<env_grid<CfgV7>(n), CfgV7::threads, CfgV7::cta_smem, s>>>(a); break;
    case 3: launch_sorted<SortedB>(a, s); break;
    default: { int prc = launch_partitioned(a, s); if (prc) return prc; } break;
  }
  return cuda_rc(cudaGetLastError(), "bgym_step launch");
}

int bgym_action_mask(const BgymState* state, uint64_t* mask, int64_t n, void* stream) {
  if (n < 0 || !state || !mask) return set_err(BGYM_E_ARG, "bgym_action_mask: bad arguments");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  int grid = (int)((n + 255) / 256 < (long long)g_sm_count * 8 ? (n + 255) / 256 : (long long)g_sm_count * 8);
  action_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(state), mask, n);
  return cuda_rc(cudaGetLastError(), "bgym_action_mask launch");
}

int bgym_sample_actions(const BgymObs* obs, int32_t* actions, uint32_t seed, uint64_t step, int64_t n, void* stream) {
  if (n < 0 || !obs || !actions) return set_err(BGYM_E_ARG, "bgym_sample_actions: bad arguments");
  if (n == 0) return 0;
  int rc = ensure_device_setup();
  if (rc) return rc;
  int grid = (int)((n + 255) / 256 < (long long)g_sm_count * 8 ? (n + 255) / 256 : (long long)g_sm_count * 8);
  sample_actions_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(obs), actions, seed, step, n);
  return cuda_rc(cudaGetLastError(), "bgym_sample_actions launch");
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
  if (!mods8 && !n_cards && !jokers8 && !levels12 && !ctx)
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

// ---- host-buffer handle API ---------------------------------------------------------------------
struct BgymVec {
  int64_t n; int device; cudaStream_t stream;
  uint8_t *d_state, *d_obs, *d_term, *d_trunc, *d_decks; double* d_reward; BgymInfo* d_info; int32_t* d_actions;
  uint32_t* d_seeds; BgymDraws* d_draws;
  // pinned staging
  uint8_t *h_obs, *h_term, *h_trunc, *h_decks; double* h_reward; BgymInfo* h_info; int32_t* h_actions; uint32_t* h_seeds;
  BgymDraws* h_draws;
};

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return cuda_rc(_e, #x); } while (0)

int bgym_vec_create(BgymVec** out, int64_t n, int device) {
  if (!out || n <= 0) return set_err(BGYM_E_ARG, "bgym_vec_create: bad arguments");
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) return set_err(BGYM_E_NODEV, "bgym_vec_create: no CUDA device");
  if (device < 0 || device >= cnt) return set_err(BGYM_E_ARG, "bgym_vec_create: bad device index");
  CK(cudaSetDevice(device));
  BgymVec* v = new BgymVec();
  memset(v, 0, sizeof *v);
  v->n = n; v->device = device;
  CK(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&v->d_state, n * BGYM_STATE_BYTES)); CK(cudaMalloc(&v->d_obs, n * BGYM_OBS_BYTES));
  CK(cudaMalloc(&v->d_term, n)); CK(cudaMalloc(&v->d_trunc, n)); CK(cudaMalloc(&v->d_decks, n * 52));
  CK(cudaMalloc(&v->d_reward, n * 8)); CK(cudaMalloc(&v->d_info, n * BGYM_INFO_BYTES)); CK(cudaMalloc(&v->d_actions, n * 4));
  CK(cudaMalloc(&v->d_seeds, n * 4)); CK(cudaMalloc(&v->d_draws, n * BGYM_DRAWS_BYTES));
  CK(cudaMallocHost(&v->h_obs, n * BGYM_OBS_BYTES)); CK(cudaMallocHost(&v->h_term, n)); CK(cudaMallocHost(&v->h_trunc, n));
  CK(cudaMallocHost(&v->h_decks, n * 52)); CK(cudaMallocHost(&v->h_reward, n * 8)); CK(cudaMallocHost(&v->h_info, n * BGYM_INFO_BYTES));
  CK(cudaMallocHost(&v->h_actions, n * 4)); CK(cudaMallocHost(&v->h_seeds, n * 4)); CK(cudaMallocHost(&v->h_draws, n * BGYM_DRAWS_BYTES));
  CK(cudaMemsetAsync(v->d_state, 0, n * BGYM_STATE_BYTES, v->stream));
  *out = v;
  return 0;
}

int bgym_vec_destroy(BgymVec* v) {
  if (!v) return 0;
  cudaSetDevice(v->device);
  cudaStreamSynchronize(v->stream);
  cudaFree(v->d_state); cudaFree(v->d_obs); cudaFree(v->d_term); cudaFree(v->d_trunc); cudaFree(v->d_decks);
  cudaFree(v->d_reward); cudaFree(v->d_info); cudaFree(v->d_actions); cudaFree(v->d_seeds); cudaFree(v->d_draws);
  cudaFreeHost(v->h_obs); cudaFreeHost(v->h_term); cudaFreeHost(v->h_trunc); cudaFreeHost(v->h_decks);
  cudaFreeHost(v->h_reward); cudaFreeHost(v->h_info); cudaFreeHost(v->h_actions); cudaFreeHost(v->h_seeds); cudaFreeHost(v->h_draws);
  cudaStreamDestroy(v->stream);
  delete v;
  return 0;
}

int bgym_vec_reset_host(BgymVec* v, const uint32_t* seeds, const uint8_t* decks52, BgymObs* obs_out) {
  if (!v || !seeds) return set_err(BGYM_E_ARG, "bgym_vec_reset_host: bad arguments");
  CK(cudaSetDevice(v->device));
  memcpy(v->h_seeds, seeds, v->n * 4);
  CK(cudaMemcpyAsync(v->d_seeds, v->h_seeds, v->n * 4, cudaMemcpyHostToDevice, v->stream));
  if (decks52) {
    memcpy(v->h_decks, decks52, v->n * 52);
    CK(cudaMemcpyAsync(v->d_decks, v->h_decks, v->n * 52, cudaMemcpyHostToDevice, v->stream));
  }
  int rc = bgym_reset(reinterpret_cast<BgymState*>(v->d_state), reinterpret_cast<BgymObs*>(v->d_obs), nullptr, v->d_seeds,
                      decks52 ? v->d_decks : nullptr, v->n, 0, v->stream);
  if (rc) return rc;
  if (obs_out) CK(cudaMemcpyAsync(v->h_obs, v->d_obs, v->n * BGYM_OBS_BYTES, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  if (obs_out) memcpy(obs_out, v->h_obs, v->n * BGYM_OBS_BYTES);
  return 0;
}

int bgym_vec_step_host(BgymVec* v, const int32_t* actions, const BgymDraws* draws, BgymObs* obs_out,
                       double* reward_out, uint8_t* terminated_out, uint8_t* truncated_out,
                       BgymInfo* info_out, int flags) {
  if (!v || !actions) return set_err(BGYM_E_ARG, "bgym_vec_step_host: bad arguments");
  CK(cudaSetDevice(v->device));
  memcpy(v->h_actions, actions, v->n * 4);
  CK(cudaMemcpyAsync(v->d_actions, v->h_actions, v->n * 4, cudaMemcpyHostToDevice, v->stream));
  if (draws) {
    memcpy(v->h_draws, draws, v->n * BGYM_DRAWS_BYTES);
    CK(cudaMemcpyAsync(v->d_draws, v->h_draws, v->n * BGYM_DRAWS_BYTES, cudaMemcpyHostToDevice, v->stream));
  }
  int rc = bgym_step(reinterpret_cast<BgymState*>(v->d_state), v->d_actions, draws ? v->d_draws : nullptr,
                     reinterpret_cast<BgymObs*>(v->d_obs), v->d_reward, v->d_term, v->d_trunc, v->d_info, v->n,
                     flags & ~BGYM_FLAG_NO_OBS, v->stream);
  if (rc) return rc;
  if (obs_out) CK(cudaMemcpyAsync(v->h_obs, v->d_obs, v->n * BGYM_OBS_BYTES, cudaMemcpyDeviceToHost, v->stream));
  if (reward_out) CK(cudaMemcpyAsync(v->h_reward, v->d_reward, v->n * 8, cudaMemcpyDeviceToHost, v->stream));
  if (terminated_out) CK(cudaMemcpyAsync(v->h_term, v->d_term, v->n, cudaMemcpyDeviceToHost, v->stream));
  if (truncated_out) CK(cudaMemcpyAsync(v->h_trunc, v->d_trunc, v->n, cudaMemcpyDeviceToHost, v->stream));
  if (info_out) CK(cudaMemcpyAsync(v->h_info, v->d_info, v->n * BGYM_INFO_BYTES, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  if (obs_out) memcpy(obs_out, v->h_obs, v->n * BGYM_OBS_BYTES);
  if (reward_out) memcpy(reward_out, v->h_reward, v->n * 8);
  if (terminated_out) memcpy(terminated_out, v->h_term, v->n);
  if (truncated_out) memcpy(truncated_out, v->h_trunc, v->n);
  if (info_out) memcpy(info_out, v->h_info, v->n * BGYM_INFO_BYTES);
  return 0;
}

int bgym_vec_pointers(BgymVec* v, void** state, void** obs, void** reward, void** terminated) {
  if (!v) return set_err(BGYM_E_ARG, "bgym_vec_pointers: null handle");
  if (state) *state = v->d_state;
  if (obs) *obs = v->d_obs;
  if (reward) *reward = v->d_reward;
  if (terminated) *terminated = v->d_term;
  return 0;
}

int bgym_vec_get_state(BgymVec* v, BgymState* host_out) {
  if (!v || !host_out) return set_err(BGYM_E_ARG, "bgym_vec_get_state: bad arguments");
  CK(cudaSetDevice(v->device));
  CK(cudaMemcpyAsync(host_out, v->d_state, v->n * BGYM_STATE_BYTES, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  return 0;
}

int bgym_vec_set_state(BgymVec* v, const BgymState* host_in) {
  if (!v || !host_in) return set_err(BGYM_E_ARG, "bgym_vec_set_state: bad arguments");
  CK(cudaSetDevice(v->device));
  CK(cudaMemcpyAsync(v->d_state, host_in, v->n * BGYM_STATE_BYTES, cudaMemcpyHostToDevice, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  return 0;
}

}  // extern "C"
