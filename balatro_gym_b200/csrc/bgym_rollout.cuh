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
// policy_first_layer_kernel: the FIRST Linear + ReLU of the extractor's three sub-nets (train_balatro_agent.py:52-69:
// hand_net[0] 416 -> 256, joker_net[0] 10 -> 128, game_state_net[0] 21 -> 64) computed straight from the observation
// records.  The hand block of the input is an 8-hot vector (:86-93), so hand_net[0] is a SUM OF EIGHT ROWS of the
// transposed weight, one per hand slot; featurize_kernel + a 416-wide GEMM wrote and re-read 896 B of mostly zeros
// per env for it.  One persistent CTA per SM keeps all three transposed weight matrices in shared memory (213 KB of
// bf16 + 5 KB + biases); a warp serves one env at a time: lane l owns outputs [8l, 8l+8) of the hand block (eight
// conflict-free LDS.128), [4l, 4l+4) of the joker block, [2l, 2l+2) of the game block; fp32 accumulation, bf16 out
// (the operands are the bf16 values the GEMM path multiplies: one-hot ones, joker ids, bf16-rounded game scalars).
// ------------------------------------------------------------------------------------------------
constexpr int FL_HAND = 256, FL_JOKER = 128, FL_GAME = 64, FL_OUT = FL_HAND + FL_JOKER + FL_GAME;   // 448
constexpr int FL_WARPS = 8;
constexpr int FL_SMEM = (416 * FL_HAND + 10 * FL_JOKER + 21 * FL_GAME) * 2 + FL_OUT * 4;           // 220 032 B

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void bf16x2_add(float& a0, float& a1, uint32_t w, float scale) {
  a0 += scale * __uint_as_float(w << 16);
  a1 += scale * __uint_as_float(w & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(fmaxf(a, 0.0f), fmaxf(b, 0.0f));
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __launch_bounds__(FL_WARPS * 32, 1) policy_first_layer_kernel(const uint8_t* __restrict__ obs,
    const __nv_bfloat16* __restrict__ wt_hand, const __nv_bfloat16* __restrict__ wt_joker, const __nv_bfloat16* __restrict__ wt_game,
    const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, long long n) {
  extern __shared__ __align__(16) uint8_t fl_smem[];
  uint8_t* s_hand = fl_smem;                                   // [416][256] bf16
  uint8_t* s_joker = s_hand + 416 * FL_HAND * 2;                // [10][128]
  uint8_t* s_game = s_joker + 10 * FL_JOKER * 2;                // [21][64]
  float* s_bias = reinterpret_cast<float*>(s_game + 21 * FL_GAME * 2);   // [448]
  for (int i = threadIdx.x; i < 416 * FL_HAND * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_hand)[i] = __ldg(reinterpret_cast<const uint4*>(wt_hand) + i);
  for (int i = threadIdx.x; i < 10 * FL_JOKER * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_joker)[i] = __ldg(reinterpret_cast<const uint4*>(wt_joker) + i);
  for (int i = threadIdx.x; i < 21 * FL_GAME * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_game)[i] = __ldg(reinterpret_cast<const uint4*>(wt_game) + i);
  for (int i = threadIdx.x; i < FL_OUT; i += blockDim.x) s_bias[i] = bias[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long env = (long long)blockIdx.x * FL_WARPS + warp; env < n; env += (long long)gridDim.x * FL_WARPS) {
    const uint8_t* o = obs + env * BGYM_OBS_BYTES;
    // ---- hand block: bias + one weight row per occupied hand slot ----
    float h[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(s_bias + 8 * lane), b1 = *reinterpret_cast<const float4*>(s_bias + 8 * lane + 4);
      h[0] = b0.x; h[1] = b0.y; h[2] = b0.z; h[3] = b0.w; h[4] = b1.x; h[5] = b1.y; h[6] = b1.z; h[7] = b1.w;
    }
    const unsigned long long hand = __ldg(reinterpret_cast<const unsigned long long*>(o));     // 8 x int8, -1 = empty
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int c = (int)(int8_t)(hand >> (8 * i));
      if (c >= 0 && c < 52) {            // warp-uniform: every lane looks at the same env
        const uint4 w = lds128(s_hand + ((i * 52 + c) * FL_HAND + 8 * lane) * 2);
        bf16x2_add(h[0], h[1], w.x, 1.0f); bf16x2_add(h[2], h[3], w.y, 1.0f);
        bf16x2_add(h[4], h[5], w.z, 1.0f); bf16x2_add(h[6], h[7], w.w, 1.0f);
      }
    }
    // ---- joker block: joker_ids[k] as a float times row k ----
    float j[4];
    {
      const float4 b = *reinterpret_cast<const float4*>(s_bias + FL_HAND + 4 * lane);
      j[0] = b.x; j[1] = b.y; j[2] = b.z; j[3] = b.w;
    }
#pragma unroll
    for (int k = 0; k < 10; k++) {
      const float id = i16f(o, 64 + 2 * k);
      if (id != 0.0f) {
        const uint2 w = *reinterpret_cast<const uint2*>(s_joker + (k * FL_JOKER + 4 * lane) * 2);
        bf16x2_add(j[0], j[1], w.x, id); bf16x2_add(j[2], j[3], w.y, id);
      }
    }
    // ---- game block: the 21 scaled scalars (feature_scalars) times their rows ----
    float g0 = s_bias[FL_HAND + FL_JOKER + 2 * lane], g1 = s_bias[FL_HAND + FL_JOKER + 2 * lane + 1];
    float v1[8], v2[8], v3[8];
    feature_scalars(o, 1, v1); feature_scalars(o, 2, v2); feature_scalars(o, 3, v3);
#pragma unroll
    for (int k = 0; k < 21; k++) {
      const float f = bf16_round(k < 6 ? v1[2 + k] : (k < 14 ? v2[k - 6] : v3[k - 14]));
      const uint32_t w = *reinterpret_cast<const uint32_t*>(s_game + (k * FL_GAME + 2 * lane) * 2);
      bf16x2_add(g0, g1, w, f);
    }
    // ---- ReLU, bf16, coalesced stores: [hand 256 | joker 128 | game 64] ----
    __nv_bfloat16* row = out + env * FL_OUT;
    uint4 q;
    q.x = relu_pack_bf16x2(h[0], h[1]); q.y = relu_pack_bf16x2(h[2], h[3]);
    q.z = relu_pack_bf16x2(h[4], h[5]); q.w = relu_pack_bf16x2(h[6], h[7]);
    reinterpret_cast<uint4*>(row)[lane] = q;
    reinterpret_cast<uint2*>(row + FL_HAND)[lane] = make_uint2(relu_pack_bf16x2(j[0], j[1]), relu_pack_bf16x2(j[2], j[3]));
    reinterpret_cast<uint32_t*>(row + FL_HAND + FL_JOKER)[lane] = relu_pack_bf16x2(g0, g1);
  }
}

// ------------------------------------------------------------------------------------------------
// masked categorical: one thread per env, the 60 logits of its row in registers (15 x 16-byte loads; the
// two halves of every 32-byte sector are consumed by consecutive loads, so L1 absorbs the 240-byte row
// stride).  Three passes over registers: max, sum of exp (+ entropy numerator), inverse-CDF scan.
// (A 16-lanes-per-env version with half-warp shuffles was 4x the instructions per env: one Philox call per
// two envs instead of one per 32, and reductions instead of serial adds.)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) masked_sample_kernel(const T* __restrict__ logits, const uint8_t* __restrict__ mask_words,
                                                            long long mask_stride, const float* __restrict__ uniforms, uint32_t seed,
                                                            unsigned long long step, long long env_offset,
                                                            int32_t* __restrict__ actions, float* __restrict__ logp,
                                                            float* __restrict__ entropy, long long n) {
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  unsigned long long mask = *reinterpret_cast<const unsigned long long*>(mask_words + env * mask_stride);
  mask &= (1ull << BGYM_NUM_ACTIONS) - 1;
  float x[BGYM_NUM_ACTIONS];
  if constexpr (sizeof(T) == 4) {
    const float4* row = reinterpret_cast<const float4*>(logits + env * BGYM_NUM_ACTIONS);
#pragma unroll
    for (int q = 0; q < BGYM_NUM_ACTIONS / 4; q++) {
      const float4 v = row[q];
      x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
  } else {
    const uint2* row = reinterpret_cast<const uint2*>(logits + env * BGYM_NUM_ACTIONS);
#pragma unroll
    for (int q = 0; q < BGYM_NUM_ACTIONS / 4; q++) {
      uint2 v = row[q];
      const __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x), b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
      x[4 * q] = __low2float(a); x[4 * q + 1] = __high2float(a); x[4 * q + 2] = __low2float(b); x[4 * q + 3] = __high2float(b);
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < BGYM_NUM_ACTIONS; k++) if ((mask >> k) & 1ull) m = fmaxf(m, x[k]);
  float S = 0.0f, SX = 0.0f;
#pragma unroll
  for (int k = 0; k < BGYM_NUM_ACTIONS; k++) {
    const bool legal = (mask >> k) & 1ull;
    const float d = x[k] - m;
    const float e = legal ? expf(d) : 0.0f;
    x[k] = e;                       // the row becomes the unnormalised probabilities
    S += e;
    SX += legal ? e * d : 0.0f;
  }
  float u;
  if (uniforms) u = uniforms[env];
  else {
    const unsigned long long ge = (unsigned long long)(env + env_offset);
    const uint4 w = philox4x32_10((uint32_t)ge, (uint32_t)(ge >> 32), (uint32_t)step, (uint32_t)(step >> 32), seed, BGYM_SAMPLE_KEY1);
    u = (float)(w.x >> 8) * (1.0f / 16777216.0f);
  }
  const float target = u * S;
  float run = 0.0f;
  int a = -1;
#pragma unroll
  for (int k = 0; k < BGYM_NUM_ACTIONS; k++) {
    run += x[k];
    a = (a < 0 && ((mask >> k) & 1ull) && run > target) ? k : a;
  }
  if (a < 0) a = mask ? 63 - __clzll((long long)mask) : 0;   // u * S rounded up to the total: last legal action
  const bool any = mask != 0ull;
  const float logS = logf(S);
  const float xa = (float)logits[env * BGYM_NUM_ACTIONS + a];   // the chosen logit again (an L1 hit)
  actions[env] = a;
  logp[env] = any ? (xa - m) - logS : 0.0f;
  if (entropy) entropy[env] = any ? logS - SX / S : 0.0f;
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
