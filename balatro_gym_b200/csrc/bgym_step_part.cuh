// bgym_step_part.cuh — the category-partitioned step.
//
// Why it is split (measured on B200, profiles/r01_ncu_summary_history.md and r02_ncu_summary.md): with every action
// category compiled into one kernel and one env per lane, the step was bound by instruction fetch and by divergence
// (84 KB of SASS walked by warps whose lanes sit on different paths, ~11 of 32 lanes active), while a converged
// select-only step already ran at ~85 % of the HBM roofline.  Round 1 split it into a main pass and three gather
// passes (PLAY / DISCARD / everything else); ncu then showed the "everything else" pass running with 7.8 active
// lanes per instruction and the rare long paths (round advance + shop generation, autoreset, consumables) with 1-5.
// So the unit of work is now a TILE OF 32 ENVS THAT TAKE THE SAME PATH:
//
//   main pass      (every env)   reads ONLY the 32-byte toggle records (BgymTog: the selection, phase, and what the PLAY-phase
//                                mask needs), fully handles SELECT toggles (~75 % of random-legal steps) and rejected actions
//                                of envs in PLAY phase — writing back the toggle record and the 16-byte selection record
//                                (BgymSel: selected_cards + mask word), 94 B per env in all — and appends every other env to
//                                ONE OF SEVEN lists by the path its action takes (staged per warp, one atomic per run of
//                                ~100).  Hot, cold and observation records are never touched.
//   level-1 pass   (7 launches)  one kernel per list, on forked streams by a launch plan (bgym_step: long-tile lists head the
//                                streams, short-tile lists are chained behind them): each lane serves ONE listed env — hot +
//                                toggle record lifted into registers, cold record read in place —, runs that list's path (step_env<CATS> compiles only the list's
//                                branches), pushes hot (+ cold) + observation back.  Two long paths are NOT run here
//                                but handed on, again as lists: the round advance of a hand that beat the blind, and
//                                the in-place reset of a terminated env.
//   level-2 pass   (1 launch)    env_step_level_kernel<2>: tiles of envs that are reset (32 fresh decks shuffled side by side)
//                                and tiles of envs that advance a round (continuing the step's draw sequence where level 1
//                                left it).
//   Inside a tile, work that one lane would do alone is done by the warp (Immolate's deck compaction, the in-tile
//   autoreset) or by all lanes that need it at one converged point (a played hand's rolls, the consumables' bookkeeping).
//
//   small slabs (n <= 65536)     one launch of the gather tile code over all envs with every category compiled in and
//                                nothing deferred (env_step_small_kernel): such a step is launch- and latency-bound
//
// The launches are stream ordered and every env is owned by exactly one tile per level, so results do not depend on
// list order.
#pragma once
#include "bgym_env.cuh"

namespace bgym {

// lists of deferred env indices.  Level 1 (filled by the main pass): heavy tiles first, so that the persistent
// warps of the level-1 pass end on the cheap ones.  Level 2 (filled by the level-1 pass).
enum { L_PLAY = 0, L_CONS, L_GEN, L_MISC, L_DISCARD, L_SHOP, L_BLIND, L_ADVANCE, L_RESET, N_LISTS };
constexpr int N_LISTS_L1 = 7;

// list of an env the main pass does not serve itself, from what the hot record alone tells (PLAY-phase legality is
// complete there; in SHOP phase affordability needs the cold record — the tile re-checks the mask), -1 = served here
__device__ __forceinline__ int route_env(const Hot& h, int action, uint64_t play_mask, bool guard) {
  if (guard) return L_MISC;
  if (h.phase == BGYM_PHASE_PLAY) {
    if (action < 0 || action >= BGYM_NUM_ACTIONS || !((play_mask >> action) & 1)) return -1;   // rejected: no state change
    if (action == BGYM_A_PLAY_HAND) return L_PLAY;
    if (action == BGYM_A_DISCARD) return L_DISCARD;
    return action < BGYM_A_USE_CONS_BASE ? -1 : L_CONS;
  }
  if (h.phase == BGYM_PHASE_SHOP) {
    if (action == BGYM_A_SHOP_REROLL) return L_GEN;
    if (action == BGYM_A_SHOP_END) return L_BLIND;      // deals a hand, like a blind selection
    return (action >= BGYM_A_SHOP_BUY_BASE && action < BGYM_A_SELL_JOKER_BASE + 5) ? L_SHOP : L_MISC;
  }
  if (h.phase == BGYM_PHASE_BLIND_SELECT) {
    if (action == BGYM_A_SKIP_BLIND) return L_GEN;
    return (action >= BGYM_A_SELECT_BLIND_BASE && action < BGYM_A_SKIP_BLIND) ? L_BLIND : L_MISC;
  }
  return L_MISC;
}

__device__ __forceinline__ void write_step_outputs(const StepArgs& a, long long e, double reward, int terminated,
                                                   const StepInfo& info) {
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

// uniform legal action of the fused random policy: Philox keyed (seed, policy key), counter = steps
// taken in the episode
__device__ __forceinline__ int policy_action(const Hot& h, uint64_t mask) {
  int cnt = __popcll(mask);
  if (!cnt) return 0;
  uint4 w = philox4x32_10(h.ep_len, 0, 0, 0, h.rng_seed, BGYM_POLICY_KEY1);
  int k = (int)__umulhi(w.x, (uint32_t)cnt);
  return select_bit64(mask, k);
}

// ------------------------------------------------------------------------------------------------
// main pass: every env, on the toggle records alone
// ------------------------------------------------------------------------------------------------
// A card toggle (and a rejected action) needs nothing but the env's 32-byte toggle record (BgymTog, include/bgym.h): the
// pass reads tog[e] and actions[e], and for the envs it serves writes tog[e] (the selection), sel[e] (the two
// observation fields a toggle changes: selected_cards and the legal-action word), reward and flags —
// 36 B read + 58 B written per env, all of it coalesced (a warp's 32 toggle records are 1 KB contiguous) — where the
// round-1/2 pass staged 144-byte hot records through shared memory and rewrote whole 176-byte observation records
// (350 B per env).  The 176-byte observation record of a served env is NOT rewritten: none of its other fields changed.
// No shared-memory staging of records, no bulk copies: each lane holds its record in eight registers, the next tile's
// loads are issued before the current tile is served.
// Deferred env indices are staged per warp in shared memory and appended to the device lists in
// runs of ~100 (one atomic per run, PART_STAGE).  One atomic per tile and list kept a single L2 slice busy for
// most of the pass (~10^5 same-sector atomics per 2^20 envs); the counters also sit 128 B apart.
constexpr int MAIN_WARPS = 8;
#ifndef BGYM_MAIN_CTAS
#define BGYM_MAIN_CTAS 3
#endif
constexpr int MAIN_CTAS_PER_SM = BGYM_MAIN_CTAS;
#ifndef BGYM_PART_STAGE
#define BGYM_PART_STAGE 128
#endif
constexpr int PART_STAGE = BGYM_PART_STAGE;   // entries per staged list; a list is flushed when a tile could overflow it
constexpr int PART_CTR_STRIDE = 32;   // ints between list counters
constexpr int MAIN_CTA_SMEM = MAIN_WARPS * N_LISTS_L1 * PART_STAGE * 4;   // static shared memory: the staged lists

__global__ void __launch_bounds__(MAIN_WARPS * 32, MAIN_CTAS_PER_SM) env_step_main_kernel(const __grid_constant__ StepArgs a) {
  __shared__ int stage_all[MAIN_WARPS * N_LISTS_L1 * PART_STAGE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* stage_list = stage_all + warp * N_LISTS_L1 * PART_STAGE;
  int staged[N_LISTS_L1];   // warp-uniform fill of the staging lists
#pragma unroll
  for (int c = 0; c < N_LISTS_L1; c++) staged[c] = 0;
  auto flush_list = [&](int c, int cnt) {   // whole warp
    __syncwarp();
    int basei = 0;
    if (lane == 0) basei = atomicAdd(a.part_counters + c * PART_CTR_STRIDE, cnt);
    basei = __shfl_sync(0xffffffffu, basei, 0);
    int* dst = a.part_lists + (long long)c * a.part_cap + basei;
    for (int i = lane; i < cnt; i += 32) dst[i] = stage_list[c * PART_STAGE + i];
    __syncwarp();
  };
  const long long n_tiles = (a.n + 31) >> 5;
  const long long warp_gid = (long long)blockIdx.x * MAIN_WARPS + warp;
  const long long warp_cnt = (long long)gridDim.x * MAIN_WARPS;
  const bool with_obs = a.sel != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  const bool fused_policy = (a.flags & BGYM_FLAG_RANDOM_POLICY) != 0;
  // software pipeline: the toggle record and the action of this lane's env in the NEXT tile are loaded one iteration ahead
  uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
  int action_next = 0;
  auto fetch = [&](long long tile) {
    const long long en = tile * 32 + lane;
    if (tile < n_tiles && en < a.n) {
      const uint4* t = reinterpret_cast<const uint4*>(a.tog + en * BGYM_TOG_BYTES);
      n0 = __ldcs(t); n1 = __ldcs(t + 1);                       // read once: streaming
      if (!fused_policy) action_next = __ldcs(a.actions + en);
    }
  };
  fetch(warp_gid);
  for (long long tile = warp_gid; tile < n_tiles; tile += warp_cnt) {
    const long long e = tile * 32 + lane;
    const bool active = e < a.n;
    const uint4 t0 = n0, t1 = n1;
    int action = action_next;
    fetch(tile + warp_cnt);

    int cat = -2;   // -2 inactive lane, -1 served here, >= 0 list
    if (active) {
      Hot h;
      unpack_tog(t0, t1, h);
      // the PLAY-phase mask needs the toggle record only; other phases are not served here
      const bool play_phase = h.phase == BGYM_PHASE_PLAY;
      const uint64_t m0 = play_phase ? action_mask(h, nullptr) : 0ull;
      if (fused_policy) {
        if (play_phase) {
          action = policy_action(h, m0);
          if (a.actions_out) a.actions_out[e] = action;
        } else {
          action = -1;     // the MISC tile has the cold record: it samples there
        }
      }
      // guard-terminated envs (:619-623) end their episode whatever the action
      const bool guard = guard_pending(h);
      cat = (fused_policy && !play_phase) ? (int)L_MISC : route_env(h, action, m0, guard);
      if (cat < 0) {
        double reward = 0.0;
        int terminated = 0;
        StepInfo info;
        step_env<CAT_SELECT, false>(h, nullptr, nullptr, action, m0, nullptr, reward, terminated, info, nullptr);
        // a rejected action changes nothing at all; a toggle changes the selection: bytes 0..15 of the toggle record
        // (the summary half is rewritten with it so that the store covers the whole 32-byte sector)
        if (info.error_code == 0) {
          uint4* t = reinterpret_cast<uint4*>(a.tog + e * BGYM_TOG_BYTES);
          t[0] = hot_chunk1(h); t[1] = t1;
          if (with_obs) *reinterpret_cast<uint4*>(a.sel + e * BGYM_SEL_BYTES) = sel_words(h, action_mask(h, nullptr));
        }
        write_step_outputs(a, e, reward, terminated, info);
      }
    }
    // defer the other envs: stage the env index in the warp's list, flush a list before a tile could overflow it
    if (__ballot_sync(0xffffffffu, cat >= 0)) {
#pragma unroll
      for (int c = 0; c < N_LISTS_L1; c++) {
        uint32_t bal = __ballot_sync(0xffffffffu, cat == c);
        if (bal) {
          if (cat == c) stage_list[c * PART_STAGE + staged[c] + __popc(bal & ((1u << lane) - 1))] = (int)e;
          staged[c] += __popc(bal);
          if (staged[c] > PART_STAGE - 32) { flush_list(c, staged[c]); staged[c] = 0; }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < N_LISTS_L1; c++) if (staged[c]) flush_list(c, staged[c]);
}

// ------------------------------------------------------------------------------------------------
// list tiles: one listed env per lane
// ------------------------------------------------------------------------------------------------
// One warp per CTA: a CTA's slot (registers) is given back when its tile ends, not when the slowest of four tiles ends
// (tile latencies within a list spread 2-3x around their mean).  Measured on the bench loop: 4 warps x 4 CTAs 0.2057 /
// 0.2085 ms per step, 2 x 8: 0.2033, 1 x 16: 0.2021, 1 x 12 (registers capped at 168 instead of 128): 0.1986 / 0.2025.
#ifndef BGYM_GATHER_WARPS
#define BGYM_GATHER_WARPS 1
#endif
constexpr int GATHER_WARPS = BGYM_GATHER_WARPS;
// shared memory of a list kernel: one cold-record slot per lane for the lists that stage it, none for the others
constexpr int GATHER_CTA_SMEM = GATHER_WARPS * 32 * BGYM_COLD_BYTES;
#ifndef BGYM_GATHER_CTAS
#define BGYM_GATHER_CTAS 12
#endif
constexpr int GATHER_CTAS_PER_SM = BGYM_GATHER_CTAS;

// append the env of every lane with `push` set to device list `l` (one atomic per call and warp)
__device__ __forceinline__ void push_list(const StepArgs& a, int l, bool push, long long e, int lane) {
  const uint32_t bal = __ballot_sync(0xffffffffu, push);
  if (!bal) return;
  int basei = 0;
  if (lane == 0) basei = atomicAdd(a.part_counters + l * PART_CTR_STRIDE, __popc(bal));
  basei = __shfl_sync(0xffffffffu, basei, 0);
  if (push) a.part_lists[(long long)l * a.part_cap + basei + __popc(bal & ((1u << lane) - 1))] = (int)e;
}

// two lists at once: both atomics are in flight together (lanes 0 and 1), one round trip instead of two
__device__ __forceinline__ void push_lists2(const StepArgs& a, int l0, bool push0, int l1, bool push1, long long e, int lane) {
  const uint32_t bal0 = __ballot_sync(0xffffffffu, push0), bal1 = __ballot_sync(0xffffffffu, push1);
  if (!(bal0 | bal1)) return;
  int basei = 0;
  if (lane == 0 && bal0) basei = atomicAdd(a.part_counters + l0 * PART_CTR_STRIDE, __popc(bal0));
  if (lane == 1 && bal1) basei = atomicAdd(a.part_counters + l1 * PART_CTR_STRIDE, __popc(bal1));
  const int base0 = __shfl_sync(0xffffffffu, basei, 0), base1 = __shfl_sync(0xffffffffu, basei, 1);
  const uint32_t below = (1u << lane) - 1;
  if (push0) a.part_lists[(long long)l0 * a.part_cap + base0 + __popc(bal0 & below)] = (int)e;
  if (push1) a.part_lists[(long long)l1 * a.part_cap + base1 + __popc(bal1 & below)] = (int)e;
}

#ifdef BGYM_PREFETCH_L1
#define BGYM_PREFETCH(p) prefetch_l1(p)
#else
#define BGYM_PREFETCH(p) prefetch_l2(p)
#endif
#ifdef BGYM_LOCKSTEP
#define BGYM_CTA_SYNC() __syncthreads()
#else
#define BGYM_CTA_SYNC()
#endif
// tile modes
enum { TM_SAMPLE = 1,      // fused random policy: envs outside PLAY phase draw their action here (the mask needs the cold record)
       TM_DEFER = 2,       // level-1 tile: round advances and resets are handed to the level-2 lists
       TM_STAGE_COLD = 4,
       TM_COLD_CLEAN = 8 };  // with TM_STAGE_COLD: the path never changes the cold record, the copy is not written back// the path reads AND rewrites the cold record (shop inventory, deck modifiers, shuffles): it works
                           // on a per-lane copy in shared memory.  In place, every load after a store to the record misses L1
                           // (global stores do not allocate there): ncu showed those kernels at 30-50 cycles per instruction.

// per-lane copy of a cold record between global and the lane's shared-memory slot (11 x 16 B; the 176-byte slot
// stride is an odd multiple of 16, so the 128-bit accesses of a warp are bank-conflict free)
__device__ __forceinline__ void cold_to_smem_async(uint8_t* slot, const uint8_t* g) {   // 11 x cp.async, no registers held
#pragma unroll
  for (int k = 0; k < BGYM_COLD_BYTES / 16; k++) cp_async16(slot + 16 * k, g + 16 * k);
}
__device__ __forceinline__ void cold_from_smem(uint8_t* g, const uint8_t* slot) {
#pragma unroll
  for (int k = 0; k < BGYM_COLD_BYTES / 16; k++) reinterpret_cast<uint4*>(g)[k] = lds128(slot + 16 * k);
}

// Which lists work on a shared-memory copy of the cold record.  Measured (profiles/r02_step_experiments.md): staging it
// for the level-1 lists that rewrite it (shop, consumables) left their latency unchanged and cost 13 us of overlap
// between the seven kernels (22 KB of shared memory per CTA); the level-2 tiles, which are nothing but cold-record
// rewrites (a shuffle; a shop generation), went from 71 to 34 us with it.
#if defined(BGYM_STAGE_ALL)
constexpr int STG_RO = TM_STAGE_COLD, STG_RW = TM_STAGE_COLD;
#elif defined(BGYM_STAGE_RW)
constexpr int STG_RO = 0, STG_RW = TM_STAGE_COLD;
#else
constexpr int STG_RO = 0, STG_RW = 0;
#endif
__host__ __device__ constexpr int list_smem_bytes(int list) { return (list >= 7 /*level 2*/ || STG_RW || STG_RO) ? GATHER_CTA_SMEM : 0; }

// cache operators of the list tiles' record traffic (ldr128 / str128, bgym_device.cuh)
#ifdef BGYM_LD_CG
constexpr bool LIST_LD_CG = true;
#else
constexpr bool LIST_LD_CG = false;
#endif
#ifdef BGYM_ST_CG
constexpr bool LIST_ST_CG = true;
#else
constexpr bool LIST_ST_CG = false;
#endif

// the whole observation of env e: the 176-byte record AND the selection record (both carry selected_cards and the mask)
// shop_changed: the record's shop chunks (shop_items / shop_costs) can differ from what they were before the step
__device__ __forceinline__ void emit_observation(const StepArgs& a, long long e, const Hot& h, const uint8_t* cold, bool shop_changed) {
  const uint64_t m = action_mask(h, cold);
  write_obs<LIST_ST_CG>(h, cold, m, a.obs + e * BGYM_OBS_BYTES);
  str128<LIST_ST_CG>(a.sel + e * BGYM_SEL_BYTES, sel_words(h, m));
  if (a.obs_dirty) a.obs_dirty[e] = (uint8_t)(BGYM_OBS_DIRTY | (shop_changed ? BGYM_OBS_DIRTY_SHOP : 0));
}
// the shop block of the observation is all zeros outside SHOP phase and is not touched by a joker sale
__device__ __forceinline__ bool shop_block_changed(int phase_before, int phase_after, int action) {
  if (phase_before != BGYM_PHASE_SHOP && phase_after != BGYM_PHASE_SHOP) return false;
  return !(phase_before == BGYM_PHASE_SHOP && phase_after == BGYM_PHASE_SHOP &&
           action >= BGYM_A_SELL_JOKER_BASE && action < BGYM_A_SELL_JOKER_BASE + 5);
}

// One tile: lane `lane` serves listed env `e`, WORKING ON THE RECORDS WHERE THEY LIE: the hot record is lifted into
// registers with nine 16-byte loads and written back the same way, the cold record (deck, shop) is read and written in
// place through L1 — a step touches a handful of its 176 bytes (eight deck entries for a discard, the deck half for a
// played hand, the shop block in the shop), and the observation goes out as eleven 16-byte stores from registers.
// What was tried before (profiles/r02_ncu_summary.md): staging hot + cold + observation tiles in shared memory with
// per-lane bulk copies (round 1: ~50 SM-cycles per bulk copy, 96 per tile, whatever the code between them did), then
// with cooperative 16-byte cp.async / store sweeps (7 M warp-instructions of address arithmetic per step, 10 KB of
// shared memory and 128 registers per warp-tile = 16 warps per SM: long-scoreboard and fetch stalls with ~4 warps per
// scheduler to hide them).  Without the staging a tile needs no shared memory at all and occupancy is set by registers.
// One KERNEL per list (env_step_list_kernel<LIST>): its own register allocation, arguments in the constant bank, no
// calls on the common path.  A single kernel walking all lists with out-of-line tile functions was measured first
// (profiles/r02_ncu_summary.md): 2 KB stack frames per thread, 2.8 M local-memory requests per step missing L1, 200 KB of
// SASS at 75 % instruction-cache hit rate — slower than round 1's divergent passes in spite of 40 % fewer instructions.
template <int CATS, int MODE>
__device__ __forceinline__ void gather_tile(const StepArgs& a, long long e, bool active, int lane, uint8_t* cold_slot) {
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  const bool fused_policy = (a.flags & BGYM_FLAG_RANDOM_POLICY) != 0;
  const bool autoreset = (a.flags & BGYM_FLAG_AUTORESET) != 0;
  constexpr bool DEFER = (MODE & TM_DEFER) != 0;
  constexpr bool STAGE = (MODE & TM_STAGE_COLD) != 0;
  uint8_t* hot = a.hot + (active ? e : 0) * BGYM_HOT_BYTES;
  uint8_t* tog = a.tog + (active ? e : 0) * BGYM_TOG_BYTES;
  uint8_t* cold_g = a.cold + (active ? e : 0) * BGYM_COLD_BYTES;
  uint8_t* cold = STAGE ? cold_slot : cold_g;
  if (STAGE) {
    __syncwarp();     // every lane is done with the previous tile's slots
    if (active) cold_to_smem_async(cold_slot, cold_g);      // in flight together with the hot record's loads below
  }
  Hot h;
  double reward = 0.0;
  int terminated = 0, snap = -1;
  StepInfo info;
  bool want_reset = false;
  uint32_t new_seed = 0;
  int action = 0;
  uint64_t m0 = 0, imm_removed = 0;
  if (active) {
#ifdef BGYM_PF_COLD
    // the cold record's lines on their way into L1 while the hot record arrives: its first touches (deck[hand[i]], the
    // hand-play count, the shop block) are otherwise dependent DRAM round trips in the middle of the tile's chain
    if (!STAGE) { prefetch_l1(cold_g); prefetch_l1(cold_g + 128); prefetch_l1(cold_g + BGYM_COLD_BYTES - 1); }
#endif
    // read once, straight from L2 (with the fused policy the main pass wrote it for PLAY-phase envs)
    action = __ldcg(a.actions + e);
    load_hot<LIST_LD_CG>(hot, tog, h);
  }
  const int phase0 = active ? h.phase : 0;
  if (STAGE) { cp_async_wait_all(); __syncwarp(); }
  if (active) m0 = action_mask(h, cold);
  BGYM_CTA_SYNC();
  if (active) {
    if ((MODE & TM_SAMPLE) && fused_policy && (h.phase != BGYM_PHASE_PLAY || !DEFER)) {
      // envs outside PLAY phase sample here, where the mask is complete; PLAY-phase envs of a level-1 tile carry
      // the action the main pass sampled (the small-slab kernel has no main pass: every env samples here)
      action = policy_action(h, m0);
      if (a.actions_out) a.actions_out[e] = action;
    }
    step_env<CATS, DEFER>(h, hot, cold, action, m0, a.draws ? a.draws + e : nullptr, reward, terminated, info, &snap, &imm_removed);
  }
  if (CATS & CAT_CONS) {     // Immolate's deck compaction, by the whole warp (bgym_env.cuh)
    immolate_compact(imm_removed, h, hot, cold, lane);
    if (imm_removed) refresh_hand_codes(h, cold);
  }
  if (active) {
    if (terminated && autoreset) {
      info.flags |= BGYM_F_AUTORESET_DONE;
      want_reset = true;
      if (!DEFER) {
        uint32_t episode = h.episode + 1;
        new_seed = next_episode_seed(h.rng_seed);
        reset_hot(h, new_seed);
        if (a.flags & BGYM_FLAG_GEN_C3) gen_hot(h, new_seed, a.flags);
        h.episode = episode;
      }
    }
  }
  BGYM_CTA_SYNC();
  if (active) write_step_outputs(a, e, reward, terminated, info);
  bool store_state = active, emit_obs = active && with_obs;
  if (DEFER) {
    // hand the long paths on: a reset needs nothing of this tile but the outputs (the level-2 tile reads seed and
    // episode from the unchanged record), an advancing env continues from the state stored here
#ifdef BGYM_PUSH2
    push_lists2(a, L_RESET, want_reset, L_ADVANCE, snap >= 0, e, lane);
#else
    push_list(a, L_RESET, want_reset, e, lane);
    push_list(a, L_ADVANCE, snap >= 0, e, lane);
#endif
    if (snap >= 0) a.part_aux[e] = (uint16_t)snap;
    store_state = active && !want_reset;
    emit_obs = emit_obs && !want_reset && snap < 0;
  } else if (autoreset) {
    autoreset_warp(want_reset, new_seed, cold, lane, (a.flags & BGYM_FLAG_GEN_C3) != 0);
  }
  BGYM_CTA_SYNC();
  if (store_state) {
    store_hot<LIST_ST_CG>(hot, tog, h);
    if (!DEFER && want_reset) hot_clear_extra(hot);
  }
  BGYM_CTA_SYNC();
  if (emit_obs) emit_observation(a, e, h, cold, (!DEFER && want_reset) || shop_block_changed(phase0, h.phase, action));
  if (STAGE) {
    __syncwarp();     // a cooperative reset (small-slab kernel) writes other lanes' slots
    if (store_state && !(MODE & TM_COLD_CLEAN)) cold_from_smem(cold_g, cold_slot);
  }
}

// level-2 tile: the round advance of a hand that beat the blind (balatro_env_2.py:914-925 -> :1326-1392), every lane
// on the same path.  The step's outputs were written by the level-1 tile; this one finishes state and observation.
__device__ __forceinline__ void advance_tile(const StepArgs& a, long long e, bool active, uint8_t* cold_slot) {
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  __syncwarp();
  if (!active) return;
  uint8_t* hot = a.hot + e * BGYM_HOT_BYTES;
  uint8_t* cold_g = a.cold + e * BGYM_COLD_BYTES;
  cold_to_smem_async(cold_slot, cold_g);
  const uint32_t snap = a.part_aux[e];
  Hot h;
  load_hot<LIST_LD_CG>(hot, a.tog + e * BGYM_TOG_BYTES, h);
  cp_async_wait_all();
  step_env_advance(h, hot, cold_slot, a.draws ? a.draws + e : nullptr, snap);
  store_hot<LIST_ST_CG>(hot, a.tog + e * BGYM_TOG_BYTES, h);
  if (with_obs) emit_observation(a, e, h, cold_slot, true);     // a round advance enters the shop
  cold_from_smem(cold_g, cold_slot);
}

// level-2 tile: in-place reset of terminated envs (balatro_env_2.py:505-558), each lane building its own fresh
// episode; nothing is loaded but the seed and the episode counter of the old record.
__device__ __forceinline__ void reset_tile(const StepArgs& a, long long e, bool active, uint8_t* cold_slot) {
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  __syncwarp();
  if (!active) return;
  uint8_t* hot = a.hot + e * BGYM_HOT_BYTES;
  const int old_phase = a.tog[e * BGYM_TOG_BYTES + 9];        // BgymTog.phase of the episode that just ended
  const uint32_t old_seed = __ldcg(reinterpret_cast<const uint32_t*>(hot + 120));
  const uint32_t episode = __ldcg(reinterpret_cast<const uint32_t*>(hot + 132)) + 1;
  const uint32_t new_seed = next_episode_seed(old_seed);
  const bool gen = (a.flags & BGYM_FLAG_GEN_C3) != 0;
  Hot h;
  reset_hot(h, new_seed);
  if (gen) gen_hot(h, new_seed, a.flags);
  h.episode = episode;
  reset_blocks_serial(cold_slot, new_seed, nullptr, gen);      // deck build + shuffle in the lane's shared-memory slot
  store_hot<LIST_ST_CG>(hot, a.tog + e * BGYM_TOG_BYTES, h);
  hot_clear_extra(hot);
  if (with_obs) emit_observation(a, e, h, cold_slot, old_phase == BGYM_PHASE_SHOP);
  cold_from_smem(a.cold + e * BGYM_COLD_BYTES, cold_slot);
}

// Small slabs (n <= BGYM_SMALL_N): ONE launch, every env served by the tile code with all action categories compiled
// in and nothing deferred.  A step of a few thousand envs is bound by launch latency and one tile's dependent chain,
// not by bandwidth or instruction fetch, so the three-launch split only adds to it: this is what makes the N = 1
// Gymnasium facade usable.
__global__ void __launch_bounds__(GATHER_WARPS * 32, GATHER_CTAS_PER_SM) env_step_small_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* cold_slot = smem + (warp * 32 + lane) * BGYM_COLD_BYTES;
  const long long n_tiles = (a.n + 31) >> 5;
  for (long long tile0 = (long long)blockIdx.x * GATHER_WARPS; tile0 < n_tiles; tile0 += (long long)gridDim.x * GATHER_WARPS) {
    const long long e = (tile0 + warp) * 32 + lane;     // CTA-uniform trip count
    const bool active = e < a.n;
    gather_tile<CAT_ALL, TM_SAMPLE | TM_STAGE_COLD>(a, active ? e : -1, active, lane, cold_slot);
  }
}

// The list passes: one kernel per list, persistent warps walk its tiles.  The level-1 kernels run concurrently (forked
// streams), the two level-2 kernels after them.
static_assert(L_PLAY == 0 && L_ADVANCE == 7 && N_LISTS == 9, "list_smem_bytes() / list_ctas() name lists by value");
// resident CTAs per SM a list kernel is compiled for (register cap).  A list pass is latency-bound — a tile is one long
// dependent chain of ~20-45 us — so the level's duration is (tiles / resident warps) x tile latency.  More resident
// warps by a lower register cap cost more in spills than they bring (measured: 5 / 6 / 8 CTAs = +3 / +19 / +42 %).
#ifndef BGYM_PLAY_CTAS
#define BGYM_PLAY_CTAS GATHER_CTAS_PER_SM
#endif
__host__ __device__ constexpr int list_ctas(int list) { return list == 0 /*L_PLAY*/ ? BGYM_PLAY_CTAS : GATHER_CTAS_PER_SM; }
// one tile of list LIST: lane `lane` serves env e (bgym_step_part.cuh head: which path each list is)
template <int LIST>
__device__ __forceinline__ void list_tile(const StepArgs& a, long long e, bool active, int lane, uint8_t* cold_slot) {
  if (LIST == L_PLAY) gather_tile<CAT_PLAY, TM_DEFER | STG_RO>(a, e, active, lane, cold_slot);
  else if (LIST == L_CONS) gather_tile<CAT_CONS, TM_DEFER | STG_RW>(a, e, active, lane, cold_slot);
  else if (LIST == L_GEN) gather_tile<CAT_GEN, TM_DEFER | STG_RW>(a, e, active, lane, cold_slot);
  else if (LIST == L_MISC) gather_tile<CAT_OTHER | CAT_SELECT, TM_DEFER | TM_SAMPLE | STG_RW>(a, e, active, lane, cold_slot);
  else if (LIST == L_DISCARD) gather_tile<CAT_DISCARD, TM_DEFER | STG_RO | TM_COLD_CLEAN>(a, e, active, lane, cold_slot);
  else if (LIST == L_SHOP) gather_tile<CAT_SHOP, TM_DEFER | STG_RW>(a, e, active, lane, cold_slot);
  else if (LIST == L_BLIND) gather_tile<CAT_BLIND | CAT_SHOP_END, TM_DEFER | STG_RO | TM_COLD_CLEAN>(a, e, active, lane, cold_slot);
  else if (LIST == L_ADVANCE) advance_tile(a, e, active, cold_slot);
  else reset_tile(a, e, active, cold_slot);
}

// ONE KERNEL PER LEVEL (BGYM_LEVEL_KERNELS=1): every tile body of the level inlined into one kernel, the warps walk the
// level's lists as one concatenated tile space, heaviest list first.  A step is then three launches on one stream — no
// forked streams, no events between the levels.  (The first single-kernel version of this round called out-of-line
// tile functions: 2 KB stack frames, see profiles/r02_step_experiments.md #1; here the bodies are inlined and the kernel
// takes the largest body's registers.)
#ifndef BGYM_L1K_CTAS
#define BGYM_L1K_CTAS GATHER_CTAS_PER_SM
#endif
__host__ __device__ constexpr int level_ctas(int level) { return level == 1 ? BGYM_L1K_CTAS : GATHER_CTAS_PER_SM; }
// out-of-line tile bodies for the level-1 kernel: each keeps its own register allocation (inlined into one kernel the seven
// bodies needed 255 registers and 2 KB of spills); the arguments stay where they are — `a` is a __grid_constant__ parameter,
// a reference to it is a pointer into the constant bank, not a copy on the stack
template <int LIST>
__device__ __noinline__ void list_tile_ool(const StepArgs& a, long long e, bool active, int lane, uint8_t* cold_slot) {
  list_tile<LIST>(a, e, active, lane, cold_slot);
}
constexpr int PART_CLAIM_CTR = L_MISC * PART_CTR_STRIDE + 24;    // spare word of the counter block: next unclaimed tile of the level
template <int LEVEL>
__global__ void __launch_bounds__(32, level_ctas(LEVEL)) env_step_level_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  uint8_t* cold_slot = smem + lane * BGYM_COLD_BYTES;
  constexpr int NL = LEVEL == 1 ? N_LISTS_L1 : 2;
  // walk order: long tiles first, so that the level ends on short ones (with the fused policy MISC holds every env outside
  // PLAY phase and goes first)
  const bool fused = LEVEL == 1 && (a.flags & BGYM_FLAG_RANDOM_POLICY) != 0;
  const int order[7] = {LEVEL == 1 ? (fused ? (int)L_MISC : (int)L_GEN) : (int)L_RESET, LEVEL == 1 ? (int)L_CONS : (int)L_ADVANCE, L_PLAY, L_SHOP, L_DISCARD,
                        L_BLIND, fused ? (int)L_GEN : (int)L_MISC};
  int cnt[NL], first[NL + 1];
  first[0] = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    cnt[k] = a.part_counters[order[k] * PART_CTR_STRIDE];
    first[k + 1] = first[k] + ((cnt[k] + 31) >> 5);
  }
  // level 1: tiles are CLAIMED (first one = blockIdx.x, the next ones from a counter; the claim for the tile after this one is
  // in flight while this one is served), so that a warp that drew short tiles takes more of them; level 2: static walk
  int t = blockIdx.x, t_next = 0;
  auto claim = [&]() { if (LEVEL == 1 && lane == 0) t_next = (int)gridDim.x + atomicAdd(a.part_counters + PART_CLAIM_CTR, 1); };
  claim();
  while (t < first[NL]) {
    int k = 0;
#pragma unroll
    for (int q = 1; q < NL; q++) k += t >= first[q];
    int list_id = 0, tile0 = 0, count = 0;
#pragma unroll
    for (int q = 0; q < NL; q++) if (k == q) { list_id = order[q]; tile0 = first[q]; count = cnt[q]; }
    const int idx = (t - tile0) * 32 + lane;
    const bool active = idx < count;
    const long long e = active ? (long long)__ldcg(a.part_lists + (long long)list_id * a.part_cap + idx) : -1;
    if (LEVEL == 1) {
      switch (list_id) {
        case L_PLAY: list_tile_ool<L_PLAY>(a, e, active, lane, cold_slot); break;
        case L_CONS: list_tile_ool<L_CONS>(a, e, active, lane, cold_slot); break;
        case L_GEN: list_tile_ool<L_GEN>(a, e, active, lane, cold_slot); break;
        case L_MISC: list_tile_ool<L_MISC>(a, e, active, lane, cold_slot); break;
        case L_DISCARD: list_tile_ool<L_DISCARD>(a, e, active, lane, cold_slot); break;
        case L_SHOP: list_tile_ool<L_SHOP>(a, e, active, lane, cold_slot); break;
        default: list_tile_ool<L_BLIND>(a, e, active, lane, cold_slot); break;
      }
      t = __shfl_sync(0xffffffffu, t_next, 0);
      claim();
    } else {
      if (list_id == L_ADVANCE) list_tile<L_ADVANCE>(a, e, active, lane, cold_slot);
      else list_tile<L_RESET>(a, e, active, lane, cold_slot);
      t += gridDim.x;
    }
  }
}

template <int LIST>
__global__ void __launch_bounds__(GATHER_WARPS * 32, list_ctas(LIST)) env_step_list_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* cold_slot = smem + (warp * 32 + lane) * BGYM_COLD_BYTES;   // only touched by the lists that stage (list_smem_bytes)
  const int count = a.part_counters[LIST * PART_CTR_STRIDE];
  const int* list = a.part_lists + (long long)LIST * a.part_cap;
  const int n_tiles = (count + 31) >> 5;
  for (int tile = blockIdx.x * GATHER_WARPS + warp; tile < n_tiles; tile += gridDim.x * GATHER_WARPS) {
    const int idx = tile * 32 + lane;
    const bool active = idx < count;
    const long long e = active ? (long long)__ldcg(list + idx) : -1;
#ifdef BGYM_TILE_CLOCK
    unsigned long long tc0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tc0));
#endif
    list_tile<LIST>(a, e, active, lane, cold_slot);
#ifdef BGYM_TILE_CLOCK
    // diagnostic build (-DBGYM_TILE_CLOCK): per list, when its first tile started, when its last tile ended, and the tile
    // durations (ns), kept in the spare words behind the list's counter; bgym_step prints them every 64 calls
    __syncwarp();
    unsigned long long tc1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tc1));
    if (lane == 0) {
      unsigned long long* d = reinterpret_cast<unsigned long long*>(a.part_counters + LIST * PART_CTR_STRIDE + 8);
      atomicMax(d + 0, ~tc0); atomicMax(d + 1, tc1); atomicAdd(d + 2, tc1 - tc0); atomicMax(d + 3, tc1 - tc0); atomicAdd(d + 4, 1ull);
    }
#endif
  }
}

}  // namespace bgym
