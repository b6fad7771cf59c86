// bgym_rollout.cuh — device side of on-device PPO rollout collection.
//
//   featurize_kernel       BgymObs records -> the dense input of BalatroFeaturesExtractor.forward
//                          (train_balatro_agent.py:84-113): 8x52 one-hot hand | joker_ids as floats |
//                          the 21 scaled game scalars, padded to 448 columns, fp32 or bf16
//   masked_sample_kernel   logits[60] + the observation's legal-action word -> action, log-prob,
//                          entropy of the masked categorical (16 lanes per env, 4 logits per lane)
//   gae_kernel             generalized advantage estimation over a [T, n] rollout, one thread per env
//                          (stable_baselines3 RolloutBuffer.compute_returns_and_advantage, which the
//                          reference trains through: train_balatro_agent.py:328-336 gamma/gae_lambda)
//
// All three are bandwidth-bound elementwise passes: coalesced 16-byte accesses, grids sized to the
// input, no shared memory.
#pragma once
#include <cuda_bf16.h>
#include "bgym_device.cuh"

namespace bgym {

constexpr int FEAT_DIM = BGYM_FEATURE_DIM;        // 448 = 416 one-hot + 10 joker ids + 21 scalars + 1 pad
constexpr int FEAT_CHUNKS = FEAT_DIM / 8;          // 56 threads per env, 8 columns each

// The 32 columns after the one-hot hand, 8 per chunk.  Scalars are scaled by multiplying with the
// fp32 reciprocal of the divisor — what torch's CUDA `tensor / python_scalar` does — so the columns
// are bit-identical to the extractor's preprocessing run on a GPU.
__device__ __forceinline__ float i8f(const uint8_t* o, int off) { return (float)*reinterpret_cast<const int8_t*>(o + off); }
__device__ __forceinline__ float i16f(const uint8_t* o, int off) { return (float)*reinterpret_cast<const int16_t*>(o + off); }
__device__ __forceinline__ void feature_scalars(const uint8_t* o, int chunk /*0..3*/, float v[8]) {
  const float r10 = 1.0f / 10.0f;
  if (chunk == 0) {            // joker_ids[0..7]
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = i16f(o, 64 + 2 * k);
  } else if (chunk == 1) {     // joker_ids[8..9], chips_scored, chips_needed, progress_ratio, money, ante, round
    v[0] = i16f(o, 80); v[1] = i16f(o, 82);
    v[2] = (float)*reinterpret_cast<const long long*>(o + 24) * (1.0f / 1e6f);
    v[3] = (float)*reinterpret_cast<const int*>(o + 44) * (1.0f / 1e5f);
    v[4] = *reinterpret_cast<const float*>(o + 36);
    v[5] = (float)*reinterpret_cast<const int*>(o + 48) * (1.0f / 100.0f);
    v[6] = i16f(o, 60) * r10;
    v[7] = i8f(o, 148) * (1.0f / 3.0f);
  } else if (chunk == 2) {     // hands_left, discards_left, hand_levels[0..5]
    v[0] = i8f(o, 149) * r10;
    v[1] = i8f(o, 150) * (1.0f / 5.0f);
#pragma unroll
    for (int k = 0; k < 6; k++) v[2 + k] = i8f(o, 134 + k) * r10;
  } else {                     // hand_levels[6..11], phase, pad
#pragma unroll
    for (int k = 0; k < 6; k++) v[k] = i8f(o, 140 + k) * r10;
    v[6] = i8f(o, 155) * (1.0f / 3.0f);
    v[7] = 0.0f;
  }
}

constexpr int FEAT_ENVS_PER_CTA = 4;   // 4 x 56 = 224 threads

template <typename T>
__global__ void __launch_bounds__(FEAT_ENVS_PER_CTA * FEAT_CHUNKS) featurize_kernel(const uint8_t* __restrict__ obs, T* __restrict__ out, long long n) {
  const int c = threadIdx.x % FEAT_CHUNKS, sub = threadIdx.x / FEAT_CHUNKS;
  for (long long env = (long long)blockIdx.x * FEAT_ENVS_PER_CTA + sub; env < n; env += (long long)gridDim.x * FEAT_ENVS_PER_CTA) {
    const uint8_t* o = obs + env * BGYM_OBS_BYTES;
    float v[8];
    const int e0 = c * 8;
    if (e0 < 416) {
      // columns e0..e0+7 of the 8x52 one-hot span at most two hand slots
      const unsigned long long hand = *reinterpret_cast<const unsigned long long*>(o);   // 8 x int8, -1 = empty
      const int slot0 = e0 / 52, card0 = e0 - slot0 * 52;
      const int h0 = (int)(int8_t)(hand >> (8 * slot0));
      const int h1 = slot0 < 7 ? (int)(int8_t)(hand >> (8 * slot0 + 8)) : -1;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int card = card0 + k;
        v[k] = (card < 52 ? h0 == card : h1 == card - 52) ? 1.0f : 0.0f;
      }
    } else {
      feature_scalars(o, c - 52, v);
    }
    if constexpr (sizeof(T) == 4) {
      float4* dst = reinterpret_cast<float4*>(out + env * FEAT_DIM + e0);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
      uint4 q;
      q.x = *reinterpret_cast<uint32_t*>(&p0); q.y = *reinterpret_cast<uint32_t*>(&p1);
      q.z = *reinterpret_cast<uint32_t*>(&p2); q.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(out + env * FEAT_DIM + e0) = q;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// masked categorical: 16 lanes per env (two envs per warp), lane s holds logits 4s..4s+3
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int half_min(int v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) masked_sample_kernel(const T* __restrict__ logits, const uint8_t* __restrict__ obs,
                                                            const float* __restrict__ uniforms, uint32_t seed,
                                                            unsigned long long step, long long env_offset,
                                                            int32_t* __restrict__ actions, float* __restrict__ logp,
                                                            float* __restrict__ entropy, long long n) {
  const long long gl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long env = gl >> 4;
  const int sub = (int)(gl & 15), lane = threadIdx.x & 31;
  const bool active = env < n;      // a whole half-warp is active or not
  unsigned long long mask = active ? *reinterpret_cast<const unsigned long long*>(obs + env * BGYM_OBS_BYTES + 160) : 0ull;
  mask &= (1ull << BGYM_NUM_ACTIONS) - 1;
  float x[4];
  bool legal[4];
  if (active && sub < 15) {
    if constexpr (sizeof(T) == 4) {
      float4 q = *reinterpret_cast<const float4*>(logits + env * BGYM_NUM_ACTIONS + sub * 4);
      x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    } else {
      uint2 q = *reinterpret_cast<const uint2*>(logits + env * BGYM_NUM_ACTIONS + sub * 4);
      __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&q.x), b = *reinterpret_cast<__nv_bfloat162*>(&q.y);
      x[0] = __low2float(a); x[1] = __high2float(a); x[2] = __low2float(b); x[3] = __high2float(b);
    }
  } else {
    x[0] = x[1] = x[2] = x[3] = 0.0f;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) legal[k] = sub < 15 && ((mask >> (sub * 4 + k)) & 1ull);
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 4; k++) if (legal[k]) m = fmaxf(m, x[k]);
  m = half_max(m);
  float e[4], s = 0.0f, sx = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    e[k] = legal[k] ? expf(x[k] - m) : 0.0f;
    s += e[k];
    sx += legal[k] ? e[k] * (x[k] - m) : 0.0f;
  }
  const float S = half_sum(s);
  const float SX = half_sum(sx);
  // inclusive scan of the per-lane sums inside the half-warp
  float incl = s;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    float up = __shfl_up_sync(0xffffffffu, incl, o);
    if ((lane & 15) >= o) incl += up;
  }
  float u;
  if (uniforms) u = active ? uniforms[env] : 0.0f;
  else {
    unsigned long long ge = (unsigned long long)(env + env_offset);
    uint4 w = philox4x32_10((uint32_t)ge, (uint32_t)(ge >> 32), (uint32_t)step, (uint32_t)(step >> 32), seed, BGYM_SAMPLE_KEY1);
    u = (float)(w.x >> 8) * (1.0f / 16777216.0f);
  }
  const float target = u * S;
  float run = incl - s;
  int cand = 64;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    run += e[k];
    if (cand == 64 && legal[k] && run > target) cand = sub * 4 + k;
  }
  int a = half_min(cand);
  if (a == 64) a = mask ? 63 - __clzll((long long)mask) : 0;   // u*S rounded up to the total: last legal action
  const float logZ = logf(S);                                   // relative to m
  const int k = a & 3;
  float mine = k == 0 ? x[0] : k == 1 ? x[1] : k == 2 ? x[2] : x[3];
  float xa = __shfl_sync(0xffffffffu, mine, (lane & 16) | (a >> 2));
  if (active && sub == 0) {
    const bool any = mask != 0ull;
    actions[env] = a;
    logp[env] = any ? (xa - m) - logZ : 0.0f;
    if (entropy) entropy[env] = any ? logZ - SX / S : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------------
// GAE(gamma, lambda): adv[t] = delta[t] + gamma*lambda*(1-done[t])*adv[t+1],
// delta[t] = r[t] + gamma*(1-done[t])*V[t+1] - V[t];  returns = adv + V
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const uint8_t* __restrict__ dones, float gamma, float lam,
                                                  float* __restrict__ adv, float* __restrict__ ret, long long T, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gae = 0.0f;
    float v_next = values[T * n + i];
    for (long long t = T - 1; t >= 0; t--) {
      const float nonterminal = dones[t * n + i] ? 0.0f : 1.0f;
      const float v = values[t * n + i];
      const float delta = rewards[t * n + i] + gamma * v_next * nonterminal - v;
      gae = delta + gamma * lam * nonterminal * gae;
      adv[t * n + i] = gae;
      ret[t * n + i] = gae + v;
      v_next = v;
    }
  }
}

}  // namespace bgym
