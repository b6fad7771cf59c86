// bgym_step_sorted.cuh — the step kernel with CTA-level path sorting.
//
// Random-legal (and trained-policy) rollouts are dominated by card-select toggles (~83 % of
// steps); PLAY / DISCARD / shop / consumable / blind actions are each a few percent.  With a fixed
// lane<->env mapping almost every warp contains every category, so every warp executes the union
// of all paths at a few active lanes each (measured: 11 of 32 lanes active on average).
//
// Here a CTA owns a tile of T = 32*WARPS consecutive envs whose records are staged in shared memory
// by ONE bulk async copy.  Threads first read the tile's actions (coalesced), classify them into
// {select, play, discard, other}, and a CTA-wide counting sort assigns envs to threads so that equal
// categories sit in adjacent lanes: most warps run only the short select path fully converged and the
// rare paths are concentrated in the last warp(s).  Any thread can serve any env of the tile because
// the records live in shared memory; per-env outputs go to the env's own slot.
#pragma once
#include "bgym_env.cuh"

namespace bgym {

template <int WARPS, int CTAS>
struct SortedCfg {
  static constexpr int warps = WARPS, ctas_per_sm = CTAS, threads = WARPS * 32, T = WARPS * 32;
  static constexpr int off_obs = T * BGYM_STATE_BYTES;
  static constexpr int off_act = off_obs + T * BGYM_OBS_BYTES;
  static constexpr int off_perm = off_act + T * 4;
  static constexpr int off_cnt = off_perm + T * 2;
  static constexpr int off_bar = (off_cnt + WARPS * 4 * 4 + 15) & ~15;
  static constexpr int cta_smem = off_bar + 16;
};

__device__ __forceinline__ int action_category(int action) {
  if (action >= BGYM_A_SELECT_BASE && action < BGYM_A_SELECT_BASE + 8) return 0;
  if (action == BGYM_A_PLAY_HAND) return 1;
  if (action == BGYM_A_DISCARD) return 2;
  return 3;
}

template <typename C>
__global__ void __launch_bounds__(C::threads, C::ctas_per_sm) env_step_sorted_kernel(StepArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int T = C::T;
  uint8_t* st = smem;
  uint8_t* ob = smem + C::off_obs;
  int* act_s = reinterpret_cast<int*>(smem + C::off_act);
  uint16_t* perm = reinterpret_cast<uint16_t*>(smem + C::off_perm);
  int* counts = reinterpret_cast<int*>(smem + C::off_cnt);   // [WARPS][4]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C::off_bar);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool with_obs = a.obs != nullptr && !(a.flags & BGYM_FLAG_NO_OBS);
  const bool fused_policy = (a.flags & BGYM_FLAG_RANDOM_POLICY) != 0;

  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();

  const long long n_tiles = (a.n + T - 1) / T;
  uint32_t parity = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * T;
    const int cnt = (int)min((long long)T, a.n - base);
    if (tid == 0) {
      mbar_arrive_expect_tx(bar, (uint32_t)cnt * BGYM_STATE_BYTES);
      bulk_g2s(st, a.state + base * BGYM_STATE_BYTES, (uint32_t)cnt * BGYM_STATE_BYTES, bar);
    }
    // ---- phase 1: this tile's actions, natural mapping (thread t <-> env t) ----
    int my_action = -1;
    if (tid < cnt) {
      if (!fused_policy) my_action = __ldg(a.actions + base + tid);
    }
    if (fused_policy) {
      // the policy needs the mask, i.e. the state: wait for the tile first
      mbar_wait(bar, parity);
      if (tid < cnt) {
        Hot hp;
        const uint8_t* prec = st + tid * BGYM_STATE_BYTES;
        unpack_hot(prec, hp);
        uint64_t m0 = action_mask(hp, prec);
        int c = __popcll(m0);
        uint4 w = philox4x32_10(hp.ep_len, 0, 0, 0, hp.rng_seed, BGYM_POLICY_KEY1);
        int k = (int)__umulhi(w.x, (uint32_t)c);
#pragma unroll 1
        for (int i = 0; i < k; i++) m0 &= m0 - 1;
        my_action = c ? __ffsll((long long)m0) - 1 : 0;
        if (a.actions_out) a.actions_out[base + tid] = my_action;
      }
    }
    act_s[tid] = my_action;
    const int cat = (tid < cnt) ? action_category(my_action) : 4;
    // ---- phase 2: CTA-wide counting sort by category ----
    uint32_t b0 = __ballot_sync(0xffffffffu, cat == 0), b1 = __ballot_sync(0xffffffffu, cat == 1);
    uint32_t b2 = __ballot_sync(0xffffffffu, cat == 2), b3 = __ballot_sync(0xffffffffu, cat == 3);
    if (lane == 0) {
      counts[warp * 4 + 0] = __popc(b0); counts[warp * 4 + 1] = __popc(b1);
      counts[warp * 4 + 2] = __popc(b2); counts[warp * 4 + 3] = __popc(b3);
    }
    __syncthreads();
    if (cat < 4) {
      uint32_t mine = cat == 0 ? b0 : cat == 1 ? b1 : cat == 2 ? b2 : b3;
      int slot = __popc(mine & ((1u << lane) - 1));
#pragma unroll
      for (int w = 0; w < C::warps; w++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
          int v = counts[w * 4 + c];
          if (c < cat || (c == cat && w < warp)) slot += v;
        }
      }
      perm[slot] = (uint16_t)tid;
    }
    __syncthreads();
    const int env = (tid < cnt) ? (int)perm[tid] : -1;
    if (!fused_policy) mbar_wait(bar, parity);
    parity ^= 1;

    // ---- phase 3: serve env `env` of the tile ----
    Hot h;
    double reward = 0.0;
    int terminated = 0;
    StepInfo info;
    bool want_reset = false;
    uint32_t new_seed = 0;
    uint8_t* rec = st + max(env, 0) * BGYM_STATE_BYTES;
    if (env >= 0) {
      unpack_hot(rec, h);
      uint64_t m0 = action_mask(h, rec);
      step_env<CAT_ALL>(h, rec, act_s[env], m0, a.draws ? a.draws + base + env : nullptr, reward, terminated, info);
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
    if (env >= 0) {
      pack_hot(rec, h);
      if (with_obs) write_obs(h, rec, action_mask(h, rec), ob + env * BGYM_OBS_BYTES);
      fence_async_smem();
      const long long e = base + env;
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
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(a.state + base * BGYM_STATE_BYTES, st, (uint32_t)cnt * BGYM_STATE_BYTES);
      if (with_obs) bulk_s2g(a.obs + base * BGYM_OBS_BYTES, ob, (uint32_t)cnt * BGYM_OBS_BYTES);
      bulk_commit();
      bulk_wait_read0();   // shared memory may be overwritten by the next tile's load after this
    }
    __syncthreads();
  }
  if (tid == 0) bulk_wait0();
}

}  // namespace bgym
