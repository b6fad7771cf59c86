// bgym_step_part.cuh — the category-partitioned step.
//
// Why it is split (measured on B200, profiles/r01_ncu_summary_history.md): with every action category
// compiled into one kernel the step was bound by instruction fetch (84 KB of SASS walked by 12
// warps at different PCs, ~11 of 32 lanes active, 18 % of HBM peak), while a converged select-only
// step already ran at ~85 % of the HBM roofline.  So one env-step is split by ACTION CATEGORY into
// kernels whose code is small and whose warps all execute the same path:
//
//   main pass   (every env)    stages ONLY the hot records of each warp's tile (one bulk copy of
//                              32 x 144 B), fully handles SELECT toggles (~83 % of random-legal
//                              steps) and never-legal action ids of envs in PLAY phase — writing back only
//                              the 16-byte chunk of the hot record a toggle changes — and appends the env
//                              index of everything else to one of three device lists (staged per warp,
//                              one atomic per run of >= 32).  The cold records are never touched.
//   gather pass (one per list) each lane pulls ONE listed env's hot + cold record with its own bulk
//                              copies, runs the category's path converged, pushes hot (+ cold) +
//                              observation back.  The three passes run concurrently.
//
//   small slabs (n <= 65536)   one launch of the gather tile code over all envs with every category compiled
//                              in (env_step_small_kernel): such a step is launch- and latency-bound
//
// The launches are stream ordered after the main pass and touch disjoint envs, so results do not
// depend on list order.
#pragma once
#include "bgym_env.cuh"

namespace bgym {

// 0 = handled by the main pass, 1 = PLAY list, 2 = DISCARD list, 3 = OTHER list
__device__ __forceinline__ int action_category_part(int action) {
  if (action == BGYM_A_PLAY_HAND) return 1;
  if (action == BGYM_A_DISCARD) return 2;
  if (action >= BGYM_A_SELECT_BASE && action < BGYM_A_SELECT_BASE + 8) return 0;
  if ((action >= BGYM_A_USE_CONS_BASE && action < BGYM_A_USE_CONS_BASE + 5) ||
      (action >= BGYM_A_SHOP_BUY_BASE && action < BGYM_A_SELL_JOKER_BASE + 5) ||
      (action >= BGYM_A_SELECT_BLIND_BASE && action <= BGYM_A_SKIP_BLIND)) return 3;
  return 0;  // ids that are never legal: rejected by the mask wherever they are handled
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
#pragma unroll 1
  for (int i = 0; i < k; i++) mask &= mask - 1;
  return __ffsll((long long)mask) - 1;
}

// ------------------------------------------------------------------------------------------------
// main pass: per-warp tiles of 32 hot records, one bulk load + two bulk stores per tile
// ------------------------------------------------------------------------------------------------
constexpr int MAIN_WARPS = 4;
constexpr int MAIN_HOT_TILE = 32 * BGYM_HOT_BYTES;   // 4608
// Deferred env indices are staged per warp in shared memory and appended to the device lists in
// runs of >= 32 (one atomic per run).  One atomic per tile and list kept a single L2 slice busy for
// most of the pass (~10^5 same-sector atomics per 2^20 envs); the counters also sit 128 B apart.
constexpr int PART_STAGE = 64;
constexpr int PART_CTR_STRIDE = 32;   // ints between list counters
// STAGES = 1: one hot buffer per warp, 4 CTAs/SM (16 warps hide each other's loads)
// STAGES = 2: the next tile's hot records are prefetched while the current tile is served, 3 CTAs/SM
template <int STAGES>
struct MainCfg {
  static constexpr int warp_smem = STAGES * MAIN_HOT_TILE + 32 * BGYM_OBS_BYTES;
  // + per-warp staging of the three deferred lists (PART_STAGE entries each), then the mbarriers
  static constexpr int cta_smem = MAIN_WARPS * warp_smem + MAIN_WARPS * 3 * PART_STAGE * 4 + 16 * MAIN_WARPS;
  static constexpr int ctas_per_sm = (227 * 1024) / cta_smem;
};

template <int STAGES>
__global__ void __launch_bounds__(MAIN_WARPS * 32, MainCfg<STAGES>::ctas_per_sm) env_step_main_kernel(StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int MAIN_WARP_SMEM = MainCfg<STAGES>::warp_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* hot_base = smem + warp * MAIN_WARP_SMEM;
  uint8_t* obs_buf = hot_base + STAGES * MAIN_HOT_TILE;
  int* stage_list = reinterpret_cast<int*>(smem + MAIN_WARPS * MAIN_WARP_SMEM) + warp * 3 * PART_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MAIN_WARPS * MAIN_WARP_SMEM + MAIN_WARPS * 3 * PART_STAGE * 4) + warp * 2;
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
  __syncwarp();
  int staged[3] = {0, 0, 0};   // warp-uniform fill of the three staging lists
  auto flush_list = [&](int c) {   // whole warp; c = 0..2
    __syncwarp();
    const int cnt = staged[c];
    int basei = 0;
    if (lane == 0) basei = atomicAdd(a.part_counters + (c + 1) * PART_CTR_STRIDE, cnt);
    basei = __shfl_sync(0xffffffffu, basei, 0);
    int* dst = a.part_lists + (long long)c * a.part_cap + basei;
    for (int i = lane; i < cnt; i += 32) dst[i] = stage_list[c * PART_STAGE + i];
    staged[c] = 0;
    __syncwarp();
  };
  const long long n_tiles = (a.n + 31) >> 5;
  const long long warp_gid = (long long)blockIdx.x * MAIN_WARPS + warp;
  const long long warp_cnt = (long long)gridDim.x * MAIN_WARPS;
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  const bool fused_policy = (a.flags & BGYM_FLAG_RANDOM_POLICY) != 0;
  uint32_t parity_bits = 0;
  int stage = 0;
  auto issue_load = [&](long long t, int st) {   // lane 0 only
    uint32_t bytes = (uint32_t)(min(32LL, a.n - t * 32) * BGYM_HOT_BYTES);
    mbar_arrive_expect_tx(&bars[st], bytes);
    bulk_g2s(hot_base + st * MAIN_HOT_TILE, a.hot + t * 32 * BGYM_HOT_BYTES, bytes, &bars[st]);
  };
  if (STAGES == 2 && lane == 0 && warp_gid < n_tiles) issue_load(warp_gid, 0);
  // the action of this lane's env in the NEXT tile is fetched one iteration ahead (its DRAM latency would
  // otherwise sit between the tile's arrival and the first use)
  int action_next = 0;
  if (!fused_policy && warp_gid * 32 + lane < a.n) action_next = __ldg(a.actions + warp_gid * 32 + lane);
  for (long long tile = warp_gid; tile < n_tiles; tile += warp_cnt, stage = (STAGES == 2) ? (stage ^ 1) : 0) {
    const long long e = tile * 32 + lane;
    const bool active = e < a.n;
    uint8_t* hot_buf = hot_base + stage * MAIN_HOT_TILE;
    uint8_t* hot = hot_buf + lane * BGYM_HOT_BYTES;
    uint8_t* obs_s = obs_buf + lane * BGYM_OBS_BYTES;
    if (lane == 0) {
      // the previous tile's bulk stores have finished READING shared memory (its hot buffer is the one
      // the next load overwrites; the obs buffer is rewritten below)
      bulk_wait_read0();
      if (STAGES == 2) { if (tile + warp_cnt < n_tiles) issue_load(tile + warp_cnt, stage ^ 1); }
      else issue_load(tile, 0);
    }
    int action = action_next;
    {
      const long long en = (tile + warp_cnt) * 32 + lane;
      if (!fused_policy && en < a.n) action_next = __ldg(a.actions + en);
    }
    __syncwarp();   // lane 0 has passed its wait: the obs buffer is free for every lane
    mbar_wait(&bars[stage], (parity_bits >> stage) & 1);
    parity_bits ^= 1u << stage;

    Hot h;
    double reward = 0.0;
    int terminated = 0;
    StepInfo info;
    bool mine = false;
    int cat = 4;
    if (active) {
      unpack_hot(hot, h);
      // the PLAY-phase mask needs the hot record only; other phases are not served here
      const bool play_phase = h.phase == BGYM_PHASE_PLAY;
      uint64_t m0 = play_phase ? action_mask(h, nullptr) : 0ull;
      if (fused_policy) {
        if (play_phase) {
          action = policy_action(h, m0);
          if (a.actions_out) a.actions_out[e] = action;
        } else {
          action = -1;     // the gather pass has the cold record: it samples there
        }
      }
      cat = action_category_part(action);
      // guard-terminated envs need a deck rebuild (cold record) -> OTHER list; so does every env
      // outside PLAY phase (its observation and mask read the shop block)
      const bool guard = h.ante > 100 || h.chips_scored > 1000000000LL;
      if (!play_phase || guard) cat = 3;
      mine = cat == 0;
      if (mine) {
        step_env<CAT_SELECT>(h, hot, nullptr, action, m0, nullptr, reward, terminated, info);
        // a card toggle changes bytes 16..31 of the hot record and nothing else (include/bgym.h): that chunk goes
        // straight from registers to its place — one 32-byte sector per env instead of the 144-byte record; a
        // rejected action changes nothing at all
        if (info.error_code == 0) *reinterpret_cast<uint4*>(a.hot + e * BGYM_HOT_BYTES + 16) = hot_chunk1(h);
        if (with_obs) write_obs(h, nullptr, action_mask(h, nullptr), obs_s);
        write_step_outputs(a, e, reward, terminated, info);
      }
    }
    // defer the other categories: stage the env index in the warp's list, flush runs of >= 32
#pragma unroll
    for (int c = 0; c < 3; c++) {
      uint32_t bal = __ballot_sync(0xffffffffu, cat == c + 1);
      if (bal) {
        if (cat == c + 1) stage_list[c * PART_STAGE + staged[c] + __popc(bal & ((1u << lane) - 1))] = (int)e;
        staged[c] += __popc(bal);
        if (staged[c] >= 32) flush_list(c);
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      uint32_t cnt = (uint32_t)min(32LL, a.n - tile * 32);
      // deferred envs' observation slots hold stale bytes; their gather pass rewrites them afterwards
      if (with_obs) bulk_s2g(a.obs + tile * 32 * BGYM_OBS_BYTES, obs_buf, cnt * BGYM_OBS_BYTES);
      bulk_commit();
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) if (staged[c]) flush_list(c);
  if (lane == 0) bulk_wait0();
}

// ------------------------------------------------------------------------------------------------
// gather pass: one listed env per lane, per-lane bulk copies of its hot and cold record
// ------------------------------------------------------------------------------------------------
constexpr int GATHER_WARPS = 4;
// the observation tile (32 x 176 B) is staged OVER the hot + cold tiles once their stores have read
// them, so a warp needs 10 KB and an SM holds BGYM_GATHER_CTAS x 4 warps (the passes are bound by
// dependent integer latency: resident warps are what buys throughput)
constexpr int GATHER_WARP_SMEM = 32 * (BGYM_HOT_BYTES + BGYM_COLD_BYTES);  // 10240 >= 32 * 176
constexpr int GATHER_CTA_SMEM = GATHER_WARPS * GATHER_WARP_SMEM + 16 * GATHER_WARPS;
#ifndef BGYM_GATHER_CTAS
#define BGYM_GATHER_CTAS 4
#endif
constexpr int GATHER_CTAS_PER_SM = BGYM_GATHER_CTAS;

// One gather tile: lane `lane` serves listed env `e` (active lanes only; n_active of them, lanes 0..n_active-1).
// hot_buf / cold_buf are the warp's 32-slot staging tiles, the observation tile is staged over them.
template <int CATS, int LIST, bool STORE_COLD>
__device__ __forceinline__ void gather_tile(const StepArgs& a, uint8_t* hot_buf, uint8_t* cold_buf, uint64_t* bar,
                                            uint32_t& parity, long long e, bool active, int n_active, int lane) {
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  const bool fused_policy = (a.flags & BGYM_FLAG_RANDOM_POLICY) != 0;
  uint8_t* hot = hot_buf + lane * BGYM_HOT_BYTES;
  uint8_t* cold = cold_buf + lane * BGYM_COLD_BYTES;
  uint8_t* obs_s = hot_buf + lane * BGYM_OBS_BYTES;     // overlay, see GATHER_WARP_SMEM
  bulk_wait_read0();    // this lane's bulk stores of the previous tile have read their slots
  __syncwarp();
  if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)n_active * (BGYM_HOT_BYTES + BGYM_COLD_BYTES));
  __syncwarp();
  if (active) {
    bulk_g2s(hot, a.hot + e * BGYM_HOT_BYTES, BGYM_HOT_BYTES, bar);
    bulk_g2s(cold, a.cold + e * BGYM_COLD_BYTES, BGYM_COLD_BYTES, bar);
  }
  // read once, straight from L2
  int action = (active && !(fused_policy && LIST >= 2)) ? __ldcg(a.actions + e) : 0;
  mbar_wait(bar, parity);
  parity ^= 1;

  Hot h;
  double reward = 0.0;
  int terminated = 0;
  StepInfo info;
  bool want_reset = false;
  uint32_t new_seed = 0;
  if (active) {
    unpack_hot(hot, h);
    uint64_t m0 = action_mask(h, cold);
    if (fused_policy && LIST >= 2) {
      // envs outside PLAY phase (and guard-terminated ones) sample here, where the mask is complete;
      // PLAY-phase envs already carry the action the main pass sampled (LIST 3 = small-slab kernel: no
      // main pass ran, every env samples here)
      if (LIST == 2 && h.phase == BGYM_PHASE_PLAY) action = __ldcg(a.actions + e);
      else { action = policy_action(h, m0); if (a.actions_out) a.actions_out[e] = action; }
    }
    // the OTHER list also receives SELECT / never-legal ids of envs the main pass does not serve
    step_env<(LIST == 2) ? (CAT_OTHER | CAT_SELECT) : CATS>(h, hot, cold, action, m0, a.draws ? a.draws + e : nullptr,
                                                          reward, terminated, info);
    if (terminated && (a.flags & BGYM_FLAG_AUTORESET)) {
      uint32_t episode = h.episode + 1;
      new_seed = next_episode_seed(h.rng_seed);
      reset_hot(h, new_seed);
      if (a.flags & BGYM_FLAG_GEN_C3) gen_hot(h, new_seed, a.flags);
      h.episode = episode;
      info.flags |= BGYM_F_AUTORESET_DONE;
      want_reset = true;
    }
  }
  if (a.flags & BGYM_FLAG_AUTORESET) autoreset_warp(want_reset, new_seed, cold, lane, (a.flags & BGYM_FLAG_GEN_C3) != 0);
  ShopObs so;
  uint64_t m1 = 0;
  if (active) {
    pack_hot(hot, h);
    if (want_reset) hot_clear_extra(hot);
    if (with_obs) { m1 = action_mask(h, cold); obs_shop_block(h, cold, so); }   // last reads of the cold slot
    write_step_outputs(a, e, reward, terminated, info);
  }
  fence_async_smem();   // every lane: a cooperative reset writes other lanes' slots
  __syncwarp();
  if (active) {
    bulk_s2g(a.hot + e * BGYM_HOT_BYTES, hot, BGYM_HOT_BYTES);
    if (STORE_COLD || want_reset) bulk_s2g(a.cold + e * BGYM_COLD_BYTES, cold, BGYM_COLD_BYTES);
    bulk_commit();
  }
  if (with_obs) {
    bulk_wait_read0();  // this lane's record stores have read shared memory ...
    __syncwarp();       // ... and so have all the others': the tile is free for the observations
    if (active) write_obs_regs(h, so, m1, obs_s);
    fence_async_smem();
    if (active) { bulk_s2g(a.obs + e * BGYM_OBS_BYTES, obs_s, BGYM_OBS_BYTES); bulk_commit(); }
  }
}

// Small slabs (n <= BGYM_SMALL_N): ONE launch, every env served by the gather tile code with all action
// categories compiled in (LIST 3 = identity list).  A step of a few thousand envs is bound by launch latency
// and one tile's dependent chain, not by bandwidth or instruction fetch, so the five-launch split only adds
// to it: this is what makes the N = 1 Gymnasium facade usable.
__global__ void __launch_bounds__(GATHER_WARPS * 32, GATHER_CTAS_PER_SM) env_step_small_kernel(StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* hot_buf = smem + warp * GATHER_WARP_SMEM;
  uint8_t* cold_buf = hot_buf + 32 * BGYM_HOT_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + GATHER_WARPS * GATHER_WARP_SMEM) + warp * 2;
  if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncwarp();
  const long long n_tiles = (a.n + 31) >> 5;
  uint32_t parity = 0;
  for (long long tile = (long long)blockIdx.x * GATHER_WARPS + warp; tile < n_tiles; tile += (long long)gridDim.x * GATHER_WARPS) {
    const long long e = tile * 32 + lane;
    const bool active = e < a.n;
    gather_tile<CAT_ALL, 3, true>(a, hot_buf, cold_buf, bar, parity, active ? e : -1, active, (int)min(32LL, a.n - tile * 32), lane);
  }
  bulk_wait0();
}

template <int CATS, int LIST, bool STORE_COLD>
__global__ void __launch_bounds__(GATHER_WARPS * 32, GATHER_CTAS_PER_SM) env_step_gather_kernel(StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* hot_buf = smem + warp * GATHER_WARP_SMEM;
  uint8_t* cold_buf = hot_buf + 32 * BGYM_HOT_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + GATHER_WARPS * GATHER_WARP_SMEM) + warp * 2;
  if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncwarp();
  const int count = a.part_counters[(LIST + 1) * PART_CTR_STRIDE];
  const int* list = a.part_lists + (long long)LIST * a.part_cap;
  const int n_tiles = (count + 31) >> 5;
  const int warp_gid = blockIdx.x * GATHER_WARPS + warp;
  const int warp_cnt = gridDim.x * GATHER_WARPS;
  uint32_t parity = 0;
  for (int tile = warp_gid; tile < n_tiles; tile += warp_cnt) {
    const int idx = tile * 32 + lane;
    const bool active = idx < count;
    const long long e = active ? (long long)list[idx] : -1;
    gather_tile<CATS, LIST, STORE_COLD>(a, hot_buf, cold_buf, bar, parity, e, active, min(32, count - tile * 32), lane);
  }
  bulk_wait0();
}

}  // namespace bgym
