// bgym_policy.cu — fused forward of the rollout policy on the 5th-generation tensor cores (sm_100a), C-ABI in
// include/bgym_policy.h.  Policy side of on-device PPO rollout collection; its own library.
//
// One CTA (128 threads, one per env of a 128-env tile, 1 CTA per SM, persistent over tiles):
//   X   128 KB of shared memory: the tile's activations, up to 512 columns of bf16, as eight K-blocks of 64 columns;
//       a K-block is 128 rows x 128 bytes, 128-byte swizzled (16-byte chunk c of row r sits at chunk c ^ (r & 7)) —
//       the canonical K-major SWIZZLE_128B operand layout of tcgen05.mma, so every layer's output is written exactly
//       where the next layer's MMA reads its A operand;
//   W   three 32 KB stages: weight tiles (<= 256 output rows x one 64-column K-block, pre-swizzled on the host into the
//       same layout) streamed from L2 with 1-D bulk async copies (cp.async.bulk, completion on an mbarrier), two tiles
//       ahead of the MMA that consumes them, across layer and tile boundaries;
//   D   the accumulators: all 512 columns of tensor memory (fp32, lane = env).
// Thread 0 issues the bulk copies and the MMAs (tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 16..256, K = 16, four per
// weight tile); tcgen05.commit hands a weight stage back to the copy ring and, after the last tile of a layer group,
// wakes all four warps, which read their 32 accumulator lanes with tcgen05.ld (32x32b.x32), add the bias, apply ReLU /
// tanh, round to bf16 and store the next operand into X (or the logits / value to global memory after the last group).
//
// Program (63 weight tiles per env tile, six epilogue groups):
//   g0  hand_net.2 256->128 | joker_net.2 128->64 | game_state_net.2 64->32   (three MMAs chains into columns 0..223)   ReLU
//   g1  combined_net.0 224(256)->512      g2  combined_net.2 512->512                                                   ReLU
//   g3  pi.0 512->256 | vf.0 512->256     g4  pi.2 256->256 | vf.2 256->256                                             tanh
//   g5  pi.4 256->60(64) | vf.4 256->1(16)                                                                              none -> global
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>

#include "bgym_policy.h"

namespace {

constexpr int TILE_M = 128;
constexpr int KB_BYTES = TILE_M * 128;            // one K-block of the activation tile: 16 KB
constexpr int X_BYTES = 8 * KB_BYTES;             // 128 KB
constexpr int STAGE_BYTES = 256 * 128;            // 32 KB: the largest weight tile
constexpr int N_STAGES = 3;
constexpr int BAR_OFFSET = X_BYTES + N_STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 128;
constexpr int IN_CHUNKS = BGYM_POLICY_IN_DIM / 8; // 56 x 16 B per input row

struct LayerDef { int n_real, k_real, n_pad, k_blocks, a_kb, col, group; };
// n_pad = MMA N (a multiple of 16; 512 is issued as two halves of 256)
constexpr LayerDef LAYERS[BGYM_POLICY_LAYERS] = {
    {128, 256, 128, 4, 0, 0, 0},   {64, 128, 64, 2, 4, 128, 0},   {32, 64, 32, 1, 6, 192, 0},
    {512, 224, 512, 4, 0, 0, 1},   {512, 512, 512, 8, 0, 0, 2},
    {256, 512, 256, 8, 0, 0, 3},   {256, 512, 256, 8, 0, 256, 3},
    {256, 256, 256, 4, 0, 0, 4},   {256, 256, 256, 4, 4, 256, 4},
    {60, 256, 64, 4, 0, 0, 5},     {1, 256, 16, 4, 4, 64, 5}};
constexpr int N_GROUPS = 6;

int build_program(BgymPolicyStep* steps, int64_t* weight_bytes) {
  int s = 0;
  int64_t off = 0;
  for (int l = 0; l < BGYM_POLICY_LAYERS; l++) {
    const LayerDef& L = LAYERS[l];
    for (int n0 = 0; n0 < L.n_pad; n0 += 256) {
      const int n = L.n_pad - n0 < 256 ? L.n_pad - n0 : 256;
      for (int kb = 0; kb < L.k_blocks; kb++) {
        BgymPolicyStep& st = steps[s++];
        st.offset = (int32_t)off; st.bytes = n * 128; st.layer = l; st.n0 = n0; st.n = n; st.kb = kb;
        st.a_kb = L.a_kb + kb; st.col = L.col + n0; st.first = kb == 0; st.last = 0; st.group = L.group; st._pad = 0;
        off += st.bytes;
      }
    }
    if (l + 1 == BGYM_POLICY_LAYERS || LAYERS[l + 1].group != L.group) steps[s - 1].last = 1;
  }
  if (weight_bytes) *weight_bytes = off;
  return s;
}
__host__ __device__ constexpr int bias_offset(int group) { return group == 0 ? 0 : 256 + 512 * (group - 1); }
constexpr int BIAS_FLOATS = 256 + 512 * (N_GROUPS - 1);

__constant__ BgymPolicyStep c_prog[BGYM_POLICY_MAX_STEPS];

// ---- PTX helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor): start address >> 4 | LBO 1 (unused with
// a swizzle) << 16 | SBO = 1024 B between 8-row groups, >> 4, << 32 | version 1 << 46 | layout type 2 << 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A and B bf16 (1 << 7, 1 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
      "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// byte address of 16-byte chunk `chunk` (0..63 over the 512 columns) of row `row` in the swizzled activation buffer
__device__ __forceinline__ uint32_t x_offset(int row, int chunk) {
  return (uint32_t)((chunk >> 3) * KB_BYTES + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

enum { ACT_RELU = 1, ACT_TANH = 2 };
// accumulator columns [0, ncols) of this thread's lane -> act(acc + bias) -> bf16 -> X columns [0, ncols)
template <int ACT>
__device__ __forceinline__ void epilogue_to_x(uint32_t taddr, int row, int ncols, const float* __restrict__ bias, uint8_t* X) {
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float x = __uint_as_float(v[8 * q + j]) + __ldg(bias + c0 + 8 * q + j);
        f[j] = ACT == ACT_RELU ? fmaxf(x, 0.0f) : tanh_fast(x);
      }
      uint4 w;
      w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]); w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(X + x_offset(row, (c0 >> 3) + q)) = w;
    }
  }
}

__global__ void __launch_bounds__(TILE_M, 1) policy_mlp_kernel(const uint4* __restrict__ act, const uint8_t* __restrict__ weights,
                                                               const float* __restrict__ bias, float* __restrict__ logits,
                                                               float* __restrict__ value, long long n, int n_steps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* X = smem;
  uint8_t* W = smem + X_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);   // [N_STAGES] weight tile landed
  uint64_t* empty = full + N_STAGES;                                 // [N_STAGES] the MMAs that read the stage are done
  uint64_t* acc_bar = empty + N_STAGES;                              // the group's accumulators are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < N_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {      // one warp allocates all 512 columns of tensor memory (1 CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes
  const int row = tid;

  const long long n_tiles = (n + TILE_M - 1) / TILE_M;
  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long total_steps = my_tiles * n_steps;
  long long produced = 0, consumed = 0;        // thread 0: weight tiles requested / handed to the tensor core
  uint32_t acc_parity = 0;
  // thread 0: request weight tiles up to (but not including) index `upto`
  auto top_up = [&](long long upto) {
    while (produced < upto && produced < total_steps) {
      const int st = (int)(produced % N_STAGES);
      if (produced >= N_STAGES) mbar_wait(&empty[st], (uint32_t)((produced / N_STAGES - 1) & 1));
      const BgymPolicyStep& ps = c_prog[produced % n_steps];
      mbar_expect_tx(&full[st], (uint32_t)ps.bytes);
      bulk_g2s(W + st * STAGE_BYTES, weights + ps.offset, (uint32_t)ps.bytes, &full[st]);
      produced++;
    }
  };
  if (tid == 0) top_up(N_STAGES);

  for (long long t = 0; t < my_tiles; t++) {
    const long long tile = blockIdx.x + t * gridDim.x;
    // ---- the tile's input: 128 rows x 448 bf16 -> K-blocks 0..6 of X (coalesced 16-byte loads, swizzled stores)
#pragma unroll 4
    for (int i = tid; i < TILE_M * IN_CHUNKS; i += TILE_M) {
      const int r = i / IN_CHUNKS, c = i - r * IN_CHUNKS;
      const long long g = tile * TILE_M + r;
      const uint4 v = g < n ? __ldg(act + g * IN_CHUNKS + c) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(X + x_offset(r, c)) = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    int s = 0;
    for (int group = 0; group < N_GROUPS; group++) {
      if (tid == 0) {
        tc_fence_after();
        for (;;) {
          const BgymPolicyStep& ps = c_prog[s];
          const int st = (int)(consumed % N_STAGES);
          mbar_wait(&full[st], (uint32_t)((consumed / N_STAGES) & 1));
          tc_fence_after();
          const uint64_t a_desc = umma_desc(smem_u32(X + ps.a_kb * KB_BYTES));
          const uint64_t b_desc = umma_desc(smem_u32(W + st * STAGE_BYTES));
          const uint32_t idesc = umma_idesc(ps.n);
#pragma unroll
          for (int k = 0; k < 4; k++)      // four K = 16 slices of the 64-column block: +32 bytes = +2 in the address field
            umma(tmem_base + (uint32_t)ps.col, a_desc + 2 * k, b_desc + 2 * k, idesc, (ps.first && k == 0) ? 0u : 1u);
          tc_commit(&empty[st]);
          consumed++;
          s++;
          const bool last = ps.last != 0;
          if (last) tc_commit(acc_bar);
          top_up(consumed + N_STAGES - 1);      // waits on the stage of the tile BEFORE the one just issued: no bubble
          if (last) break;
        }
      } else {
        // the other threads track the program position only
        while (!c_prog[s].last) s++;
        s++;
      }
      __syncwarp();      // warp 0 reconverges here: its other lanes do not spin on the barrier while thread 0 issues
      mbar_wait(acc_bar, acc_parity);
      acc_parity ^= 1;
      tc_fence_after();
      const float* gb = bias + bias_offset(group);
      if (group == 0) {
        epilogue_to_x<ACT_RELU>(taddr, row, 224, gb, X);
#pragma unroll
        for (int q = 28; q < 32; q++) *reinterpret_cast<uint4*>(X + x_offset(row, q)) = make_uint4(0, 0, 0, 0);   // columns 224..255: K padding of combined_net.0
      } else if (group <= 2) {
        epilogue_to_x<ACT_RELU>(taddr, row, 512, gb, X);
      } else if (group <= 4) {
        epilogue_to_x<ACT_TANH>(taddr, row, 512, gb, X);
      } else {
        const long long g = tile * TILE_M + row;
        uint32_t v[32];
        float4* out = reinterpret_cast<float4*>(logits + g * BGYM_POLICY_LOGITS);      // 240-byte rows: 16-byte aligned
        tmem_ld32(taddr, v);
        if (g < n) {
#pragma unroll
          for (int q = 0; q < 8; q++)
            out[q] = make_float4(__uint_as_float(v[4 * q]) + __ldg(gb + 4 * q), __uint_as_float(v[4 * q + 1]) + __ldg(gb + 4 * q + 1),
                                 __uint_as_float(v[4 * q + 2]) + __ldg(gb + 4 * q + 2), __uint_as_float(v[4 * q + 3]) + __ldg(gb + 4 * q + 3));
        }
        tmem_ld32(taddr + 32, v);
        if (g < n) {
#pragma unroll
          for (int q = 0; q < 7; q++)
            out[8 + q] = make_float4(__uint_as_float(v[4 * q]) + __ldg(gb + 32 + 4 * q), __uint_as_float(v[4 * q + 1]) + __ldg(gb + 33 + 4 * q),
                                     __uint_as_float(v[4 * q + 2]) + __ldg(gb + 34 + 4 * q), __uint_as_float(v[4 * q + 3]) + __ldg(gb + 35 + 4 * q));
        }
        tmem_ld32(taddr + 64, v);
        if (g < n) value[g] = __uint_as_float(v[0]) + __ldg(gb + 64);
      }
      // the next group's MMAs read what this epilogue wrote (generic proxy -> async proxy) and overwrite the accumulators it read
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

thread_local char g_perr[256] = "";
std::mutex g_pmu;
bool g_ready[64];
int g_sms[64];

}  // namespace

extern "C" {

const char* bgym_policy_last_error(void) { return g_perr; }

int bgym_policy_program(BgymPolicyStep* steps, int64_t* weight_bytes, int64_t* bias_floats) {
  BgymPolicyStep local[BGYM_POLICY_MAX_STEPS];
  memset(local, 0, sizeof local);
  int64_t wb = 0;
  const int n = build_program(local, &wb);
  if (steps) memcpy(steps, local, sizeof local);
  if (weight_bytes) *weight_bytes = wb;
  if (bias_floats) *bias_floats = BIAS_FLOATS;
  return n;
}

int bgym_policy_forward(const void* act, const void* weights, const float* bias, float* logits, float* value, int64_t n, void* stream) {
  if (n < 0 || !act || !weights || !bias || !logits || !value) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward: bad arguments"); return -1; }
  if (((uintptr_t)act | (uintptr_t)weights) & 15) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward: act / weights must be 16-byte aligned"); return -1; }
  if (n == 0) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward: cudaGetDevice: %s", cudaGetErrorString(e)); return (int)(e ? e : cudaErrorInvalidDevice); }
  BgymPolicyStep prog[BGYM_POLICY_MAX_STEPS];
  memset(prog, 0, sizeof prog);
  const int n_steps = build_program(prog, nullptr);
  {
    std::lock_guard<std::mutex> lock(g_pmu);
    if (!g_ready[dev]) {
      e = cudaMemcpyToSymbol(c_prog, prog, sizeof prog);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
      if (e != cudaSuccess) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward setup: %s", cudaGetErrorString(e)); return (int)e; }
      g_ready[dev] = true;
    }
  }
  const long long tiles = (n + TILE_M - 1) / TILE_M;
  const int grid = (int)(tiles < g_sms[dev] ? tiles : g_sms[dev]);
  policy_mlp_kernel<<<grid, TILE_M, SMEM_BYTES, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(act), reinterpret_cast<const uint8_t*>(weights),
                                                                        bias, logits, value, n, n_steps);
  e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

}  // extern "C"
