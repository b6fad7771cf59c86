// bgym_step_part.cuh — the category-partitioned step (default variant).
//
// Measured on B200 (profiles/): with every action category compiled into one kernel the step is
// bound by instruction fetch (84 KB of SASS walked by 12 warps at different PCs, ~11 of 32 lanes
// active), while a converged select-only step already runs at ~85 % of the HBM roofline.  So one
// env-step pass is split by ACTION CATEGORY into kernels whose code is small and whose warps all
// execute the same path:
//
//   main pass   (every env)   bulk-stages each warp's tile, fully handles SELECT toggles (~83 % of
//                             random-legal steps) + masked/guarded actions, and appends the env index
//                             of every PLAY / DISCARD / OTHER action to one of three device lists
//                             (warp-aggregated atomics);
//   gather pass (one per list) each lane pulls ONE listed env's record with its own bulk copy, runs
//                             the category's path converged, pushes state + observation back.
//
// All four launches are stream ordered; they touch disjoint envs after the main pass, so results do
// not depend on list order.
#pragma once
#include "bgym_env.cuh"

namespace bgym {

__device__ __forceinline__ int action_category_part(int action) {
  if (action == BGYM_A_PLAY_HAND) return 1;
  if (action == BGYM_A_DISCARD) return 2;
  if (action >= BGYM_A_SELECT_BASE && action < BGYM_A_SELECT_BASE + 8) return 0;
  // everything else that can be legal in some phase goes to OTHER; ids that are never legal
  // (out of range, sell-consumable, pack actions) are rejected by the mask in the main pass
  if ((action >= BGYM_A_USE_CONS_BASE && action < BGYM_A_USE_CONS_BASE + 5) ||
      (action >= BGYM_A_SHOP_BUY_BASE && action < BGYM_A_SELL_JOKER_BASE + 5) ||
      (action >= BGYM_A_SELECT_BLIND_BASE && action <= BGYM_A_SKIP_BLIND)) return 3;
  return 0;
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

// ------------------------------------------------------------------------------------------------
// main pass: per-warp tiles, one bulk load + two bulk stores per tile (as the unpartitioned kernel)
// ------------------------------------------------------------------------------------------------
constexpr int PART_WARPS = 4;
constexpr int PART_WARP_SMEM = 32 * BGYM_STATE_BYTES + 32 * BGYM_OBS_BYTES;   // 17408
constexpr int PART_CTA_SMEM = PART_WARPS * PART_WARP_SMEM + 16 * PART_WARPS;
constexpr int PART_CTAS_PER_SM = 3;

__global__ void __launch_bounds__(PART_WARPS * 32, PART_CTAS_PER_SM) env_step_main_kernel(StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* st_buf = smem + warp * PART_WARP_SMEM;
  uint8_t* obs_buf = st_buf + 32 * BGYM_STATE_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + PART_WARPS * PART_WARP_SMEM) + warp * 2;
  if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncwarp();
  const long long n_tiles = (a.n + 31) >> 5;
  const long long warp_gid = (long long)blockIdx.x * PART_WARPS + warp;
  const long long warp_cnt = (long long)gridDim.x * PART_WARPS;
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  const bool fused_policy = (a.flags & BGYM_FLAG_RANDOM_POLICY) != 0;
  uint32_t parity = 0;
  for (long long tile = warp_gid; tile < n_tiles; tile += warp_cnt) {
    const long long e = tile * 32 + lane;
    const bool active = e < a.n;
    uint8_t* rec = st_buf + lane * BGYM_STATE_BYTES;
    uint8_t* obs_s = obs_buf + lane * BGYM_OBS_BYTES;
    if (lane == 0) {
      bulk_wait_read0();   // previous tile's bulk stores have finished reading shared memory
      uint32_t bytes = (uint32_t)(min(32LL, a.n - tile * 32) * BGYM_STATE_BYTES);
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s(st_buf, a.state + tile * 32 * BGYM_STATE_BYTES, bytes, bar);
    }
    int action = 0;
    if (active && !fused_policy) action = __ldg(a.actions + e);
    mbar_wait(bar, parity);
    parity ^= 1;

    Hot h;
    double reward = 0.0;
    int terminated = 0;
    StepInfo info;
    bool want_reset = false, mine = false;
    uint32_t new_seed = 0;
    int cat = 4;
    if (active) {
      unpack_hot(rec, h);
      uint64_t m0 = action_mask(h, rec);
      if (fused_policy) {
        int cnt = __popcll(m0);
        uint4 w = philox4x32_10(h.ep_len, 0, 0, 0, h.rng_seed, BGYM_POLICY_KEY1);
        int k = (int)__umulhi(w.x, (uint32_t)cnt);
        uint64_t mm = m0;
#pragma unroll 1
        for (int i = 0; i < k; i++) mm &= mm - 1;
        action = cnt ? __ffsll((long long)mm) - 1 : 0;
        if (a.actions_out) a.actions_out[e] = action;
      }
      cat = action_category_part(action);
      mine = cat == 0;
      if (mine) {
        step_env<CAT_SELECT>(h, rec, action, m0, nullptr, reward, terminated, info);
        if (terminated && (a.flags & BGYM_FLAG_AUTORESET)) {   // guard termination only
          uint32_t episode = h.episode + 1;
          new_seed = next_episode_seed(h.rng_seed);
          reset_hot(h, new_seed);
          h.episode = episode;
          info.flags |= BGYM_F_AUTORESET_DONE;
          want_reset = true;
        }
      }
    }
    // defer the other categories: warp-aggregated append to the category's list
#pragma unroll
    for (int c = 1; c <= 3; c++) {
      uint32_t bal = __ballot_sync(0xffffffffu, cat == c);
      if (bal) {
        int leader = __ffs(bal) - 1;
        int basei = 0;
        if (lane == leader) basei = atomicAdd(a.part_counters + c, __popc(bal));
        basei = __shfl_sync(0xffffffffu, basei, leader);
        if (cat == c) a.part_lists[(long long)(c - 1) * a.part_cap + basei + __popc(bal & ((1u << lane) - 1))] = (int)e;
      }
    }
    if (a.flags & BGYM_FLAG_AUTORESET) {
      autoreset_warp(want_reset, new_seed, rec, lane);
    }
    if (mine) {
      pack_hot(rec, h);
      if (with_obs) write_obs(h, rec, action_mask(h, rec), obs_s);
      write_step_outputs(a, e, reward, terminated, info);
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      uint32_t cnt = (uint32_t)min(32LL, a.n - tile * 32);
      bulk_s2g(a.state + tile * 32 * BGYM_STATE_BYTES, st_buf, cnt * BGYM_STATE_BYTES);
      // deferred envs' observation slots hold stale bytes; their gather pass rewrites them afterwards
      if (with_obs) bulk_s2g(a.obs + tile * 32 * BGYM_OBS_BYTES, obs_buf, cnt * BGYM_OBS_BYTES);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait0();
}

// ------------------------------------------------------------------------------------------------
// gather pass: one listed env per lane, per-lane bulk copies
// ------------------------------------------------------------------------------------------------
template <int CATS, int LIST>
__global__ void __launch_bounds__(PART_WARPS * 32, PART_CTAS_PER_SM) env_step_gather_kernel(StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* st_buf = smem + warp * PART_WARP_SMEM;
  uint8_t* obs_buf = st_buf + 32 * BGYM_STATE_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + PART_WARPS * PART_WARP_SMEM) + warp * 2;
  if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncwarp();
  const int count = a.part_counters[LIST + 1];
  const int* list = a.part_lists + (long long)LIST * a.part_cap;
  const int n_tiles = (count + 31) >> 5;
  const int warp_gid = blockIdx.x * PART_WARPS + warp;
  const int warp_cnt = gridDim.x * PART_WARPS;
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  uint32_t parity = 0;
  for (int tile = warp_gid; tile < n_tiles; tile += warp_cnt) {
    const int idx = tile * 32 + lane;
    const bool active = idx < count;
    const long long e = active ? (long long)list[idx] : -1;
    uint8_t* rec = st_buf + lane * BGYM_STATE_BYTES;
    uint8_t* obs_s = obs_buf + lane * BGYM_OBS_BYTES;
    bulk_wait_read0();    // this lane's bulk stores of the previous tile have read their slots
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)min(32, count - tile * 32) * BGYM_STATE_BYTES);
    __syncwarp();
    if (active) bulk_g2s(rec, a.state + e * BGYM_STATE_BYTES, BGYM_STATE_BYTES, bar);
    int action = active ? a.actions[e] : 0;   // plain load: the fused policy wrote it in the main pass
    mbar_wait(bar, parity);
    parity ^= 1;

    Hot h;
    double reward = 0.0;
    int terminated = 0;
    StepInfo info;
    bool want_reset = false;
    uint32_t new_seed = 0;
    if (active) {
      unpack_hot(rec, h);
      uint64_t m0 = action_mask(h, rec);
      step_env<CATS>(h, rec, action, m0, a.draws ? a.draws + e : nullptr, reward, terminated, info);
      if (terminated && (a.flags & BGYM_FLAG_AUTORESET)) {
        uint32_t episode = h.episode + 1;
        new_seed = next_episode_seed(h.rng_seed);
        reset_hot(h, new_seed);
        h.episode = episode;
        info.flags |= BGYM_F_AUTORESET_DONE;
        want_reset = true;
      }
    }
    if (a.flags & BGYM_FLAG_AUTORESET) {
      autoreset_warp(want_reset, new_seed, rec, lane);
    }
    if (active) {
      pack_hot(rec, h);
      if (with_obs) write_obs(h, rec, action_mask(h, rec), obs_s);
      write_step_outputs(a, e, reward, terminated, info);
    }
    fence_async_smem();   // every lane: a cooperative reset writes other lanes' slots
    __syncwarp();
    if (active) {
      bulk_s2g(a.state + e * BGYM_STATE_BYTES, rec, BGYM_STATE_BYTES);
      if (with_obs) bulk_s2g(a.obs + e * BGYM_OBS_BYTES, obs_s, BGYM_OBS_BYTES);
      bulk_commit();
    }
  }
  bulk_wait0();
}

}  // namespace bgym
