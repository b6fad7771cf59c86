// bgym_policy.cu — fused forward of the rollout policy on the 5th-generation tensor cores (sm_100a), C-ABI in
// include/bgym_policy.h.  Policy side of on-device PPO rollout collection; its own library.
//
// One CTA (512 threads on a 128-env tile, 1 CTA per SM, persistent over tiles):
//   X   128 KB of shared memory: the tile's activations, up to 512 columns of bf16, as eight K-blocks of 64 columns;
//       a K-block is 128 rows x 128 bytes, 128-byte swizzled (16-byte chunk c of row r sits at chunk c ^ (r & 7)) —
//       the canonical K-major SWIZZLE_128B operand layout of tcgen05.mma, so every layer's output is written exactly
//       where the next layer's MMA reads its A operand;
//   W   three 32 KB stages: weight tiles (<= 256 output rows x one 64-column K-block, pre-swizzled on the host into the
//       same layout) streamed from L2 with 1-D bulk async copies (cp.async.bulk, completion on an mbarrier), two tiles
//       ahead of the MMA that consumes them, across layer and tile boundaries;
//   D   the accumulators: all 512 columns of tensor memory (fp32, lane = env).
// Warp 1 issues the bulk copies, warp 0 the MMAs (tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 16..256, K = 16, four per
// weight tile; one elected lane each, the loops themselves warp-uniform); tcgen05.commit hands a weight stage back to the copy ring and, after the last tile of a layer group,
// wakes all sixteen warps: warp w reads accumulator lanes 32 (w % 4) .. + 31 (the quarter of tensor memory a warp may
// address), columns of slice w / 4, with tcgen05.ld (32x32b.x32), adds the bias, applies ReLU / tanh, rounds to bf16 and
// stores the next operand into X (or the logits / value to global memory after the last group).  Four warps per
// scheduler matter: with one warp per scheduler the epilogues (a dependent chain of ~200 instructions per 32 columns)
// took 40 of the 86 us a tile needed.
//
// Program (72 weight tiles per env tile, seven epilogue groups):
//   g0  hand_net.0 416(448)->256 | joker_net.0 10(64)->128 | game_state_net.0 21(64)->64   from the observation record      ReLU
//   g1  hand_net.2 256->128 | joker_net.2 128->64 | game_state_net.2 64->32   (three MMA chains into columns 0..223)        ReLU
//   g2  combined_net.0 224(256)->512      g3  combined_net.2 512->512                                                       ReLU
//   g4  pi.0 512->256 | vf.0 512->256     g5  pi.2 256->256 | vf.2 256->256                                                 tanh
//   g6  pi.4 256->60(64) | vf.4 256->1(16)                                                                                  none -> global
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>

#include "bgym_policy.h"

namespace {

constexpr int TILE_M = 128;
constexpr int TILE_M_CONST = 128;
#ifndef BGYM_POLICY_THREADS
#define BGYM_POLICY_THREADS 512
#endif
constexpr int N_THREADS = BGYM_POLICY_THREADS;   // 512, or 1024 (64 registers per thread; measured 1.43 ms against 1.24)
constexpr int SLICES = N_THREADS / 128;          // column slices of an epilogue (warps per accumulator lane quarter)
constexpr int PARTS = N_THREADS / TILE_M_CONST;  // threads per env in the input stage
constexpr int KB_BYTES = TILE_M * 128;            // one K-block of the activation tile: 16 KB
constexpr int X_BYTES = 8 * KB_BYTES;             // 128 KB
constexpr int RING_BYTES = 3 * 256 * 128;         // 96 KB of weight stages: 3 x 32 KB (the largest weight tile), or 6 x 16 KB for CTA pairs
constexpr int MAX_STAGES = 6;
constexpr int BAR_OFFSET = X_BYTES + RING_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 256;
constexpr int OBS_BYTES = 176;                   // BgymObs (include/bgym.h)

struct LayerDef { int n_real, k_real, n_pad, k_blocks, a_kb, col, group; };
// n_pad = MMA N (a multiple of 16; 512 is issued as two halves of 256).  Layers 0..2 are the FIRST Linear of the three
// sub-nets, fed from the observation record: the input stage writes the 8-hot hand block (416 columns, K-blocks 0..6) and one
// more K-block (7) holding joker_ids[10] at columns 0..9 and the 21 scaled game scalars at columns 16..36.
constexpr LayerDef LAYERS[BGYM_POLICY_LAYERS] = {
    {256, 416, 256, 7, 0, 0, 0},   {128, 10, 128, 1, 7, 256, 0},  {64, 21, 64, 1, 7, 384, 0},
    {128, 256, 128, 4, 0, 0, 1},   {64, 128, 64, 2, 4, 128, 1},   {32, 64, 32, 1, 6, 192, 1},
    {512, 224, 512, 4, 0, 0, 2},   {512, 512, 512, 8, 0, 0, 3},
    {256, 512, 256, 8, 0, 0, 4},   {256, 512, 256, 8, 0, 256, 4},
    {256, 256, 256, 4, 0, 0, 5},   {256, 256, 256, 4, 4, 256, 5},
    {60, 256, 64, 4, 0, 0, 6},     {1, 256, 16, 4, 4, 64, 6}};
constexpr int N_GROUPS = 7;
constexpr int GAME_K_SHIFT = 16;      // column of K-block 7 where the game scalars start

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
__host__ __device__ constexpr int bias_offset(int group) { return 512 * group; }
constexpr int BIAS_FLOATS = 512 * N_GROUPS;

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
// one lane of a converged warp (elect.sync): the issuing warps run their loops WARP-UNIFORMLY and predicate only the issue
// itself, so that descriptors, addresses and barrier operands live in uniform registers — issued from a single divergent
// thread, every tcgen05.mma was wrapped by the compiler in a lane loop (R2UR + BRA.U.ANY) costing ~70 cycles
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// ---- CTA pairs (cta_group::2): one MMA over 256 envs, each CTA holding its 128 rows of A and HALF the rows of every weight tile
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {      // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t rank) {      // arrive on the same barrier of CTA `rank`
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void umma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
      "r"(accumulate), "r"(0u) : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor): start address >> 4 | LBO 1 (unused with
// a swizzle) << 16 | SBO = 1024 B between 8-row groups, >> 4, << 32 | version 1 << 46 | layout type 2 << 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A and B bf16 (1 << 7, 1 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc(int n, int m = TILE_M) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
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
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// byte address of 16-byte chunk `chunk` (0..63 over the 512 columns) of row `row` in the swizzled activation buffer
__device__ __forceinline__ uint32_t x_offset(int row, int chunk) {
  return (uint32_t)((chunk >> 3) * KB_BYTES + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

enum { ACT_RELU = 1, ACT_TANH = 2 };
// (lo, hi) fp32 -> packed bf16x2 of relu(.) — cvt.rn.relu.bf16x2.f32, one instruction — or of tanh(.): round to bf16, then the
// packed MUFU tanh (tanh.approx.bf16x2: two results per special-function op; with the fp32 form the two tanh epilogues were
// bound by the special-function unit, 16 lanes per clock per SM)
template <int ACT>
__device__ __forceinline__ uint32_t act_pack(float lo, float hi) {
  uint32_t r;
  if (ACT == ACT_RELU) {
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    uint32_t p;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(r) : "r"(p));
  }
  return r;
}
// 16 accumulator columns of this thread's lane, asynchronously: the registers are valid after tmem_wait16 on the same array
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
// tcgen05.wait::ld, tied to the registers it makes valid (so that no use of them is scheduled above it)
__device__ __forceinline__ void tmem_wait16(uint32_t* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]),
                 "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               : : "memory");
}
// 16 columns starting at c0: act(acc + bias) -> bf16 -> two 16-byte chunks of X
template <int ACT>
__device__ __forceinline__ void epilogue16(const uint32_t* v, int row, int c0, const float* __restrict__ bias, uint8_t* X) {
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * q)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * q + 4));
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; j++)      // fp32 bias add, then ONE instruction per pair: round to bf16 with ReLU, or round + packed tanh
      w[j] = act_pack<ACT>(__uint_as_float(v[8 * q + 2 * j]) + bb[2 * j], __uint_as_float(v[8 * q + 2 * j + 1]) + bb[2 * j + 1]);
    *reinterpret_cast<uint4*>(X + x_offset(row, (c0 >> 3) + q)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// accumulator columns [c_begin, c_end) (a multiple of 32 wide) of this thread's lane -> act(acc + bias) -> bf16 -> the same columns
// of X.  Two 16-column register buffers: the tensor-memory load of one is in flight while the other is processed.
template <int ACT>
__device__ __forceinline__ void epilogue_to_x(uint32_t taddr, int row, int c_begin, int c_end, const float* __restrict__ bias, uint8_t* X) {
  if (c_begin >= c_end) return;
  uint32_t a[16], b[16];
  tmem_ld16_async(taddr + (uint32_t)c_begin, a);
  tmem_wait16(a);
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    tmem_ld16_async(taddr + (uint32_t)(c0 + 16), b);
    epilogue16<ACT>(a, row, c0, bias, X);
    tmem_wait16(b);
    if (c0 + 32 < c_end) tmem_ld16_async(taddr + (uint32_t)(c0 + 32), a);
    epilogue16<ACT>(b, row, c0 + 16, bias, X);
    if (c0 + 32 < c_end) tmem_wait16(a);
  }
}

template <int CTAS>      // 1: every CTA on its own; 2: clusters of two CTAs issuing cta_group::2 MMAs (M = 256)
__global__ void __launch_bounds__(N_THREADS, 1) policy_mlp_kernel(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ weights,
                                                               const float* __restrict__ bias, float* __restrict__ logits,
                                                               float* __restrict__ value, long long n, int n_steps, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* X = smem;
  uint8_t* W = smem + X_BYTES;
  // a CTA of a pair holds half of every weight tile: twice the stages at half the size, i.e. a longer prefetch distance for
  // the same 96 KB (the leader learns of the peer's half through one more barrier hop)
  constexpr int N_STAGES = 3 * CTAS, STAGE_BYTES = RING_BYTES / N_STAGES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);   // [N_STAGES] weight tile landed
  uint64_t* empty = full + MAX_STAGES;                               // [N_STAGES] the MMAs that read the stage are done
  uint64_t* acc_bar = empty + MAX_STAGES;                            // the group's accumulators are complete
  uint64_t* peer_full = acc_bar + 1;                                 // [N_STAGES] (pairs, leader only) the peer's half of the tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full + MAX_STAGES);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  auto cta_sync = [&]() { if (CTAS == 2) cluster_sync(); else __syncthreads(); };    // every CTA of the MMA group
  if (tid == 0) {
    for (int i = 0; i < N_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&peer_full[i], 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {      // one warp (of each CTA of the group) allocates all 512 columns of tensor memory (1 CTA per SM)
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int quarter = warp & 3, slice = warp >> 2;                       // accumulator lanes 32 quarter .. + 31, column slice 0..SLICES-1
  // this warp's share [c_lo, c_hi) of an epilogue over ncols columns (multiples of 32)
  auto slice_lo = [&](int ncols) { const int cps = ((ncols + SLICES * 32 - 1) / (SLICES * 32)) * 32; const int lo = slice * cps; return lo < ncols ? lo : ncols; };
  auto slice_hi = [&](int ncols) { const int cps = ((ncols + SLICES * 32 - 1) / (SLICES * 32)) * 32; const int hi = slice * cps + cps; return hi < ncols ? hi : ncols; };
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
  const int row = quarter * 32 + (tid & 31);

  // a GROUP = CTAS consecutive env tiles; groups are dealt to the clusters round-robin, so both CTAs of a pair run the same
  // number of iterations (a tile past the end is processed as all-padding: it must still take part in the pair's MMAs)
  const long long n_groups_total = ((n + TILE_M - 1) / TILE_M + CTAS - 1) / CTAS;
  const long long cluster_id = blockIdx.x / CTAS, n_clusters = gridDim.x / CTAS;
  const long long my_tiles = cluster_id < n_groups_total ? (n_groups_total - cluster_id + n_clusters - 1) / n_clusters : 0;
  const long long total_steps = my_tiles * n_steps;
  const uint32_t x_base = smem_u32(X), w_base = smem_u32(W);
  // thread 0's pipeline state — small integers only: stage indices and phase bits advance by increment-and-wrap (64-bit
  // divisions and modulos in this loop cost more than the four MMAs it issues per weight tile)
  int produced = 0;                            // weight tiles requested (producer thread)
  const int total = (int)total_steps;
  int p_stage = 0, p_prog = 0, c_stage = 0;
  uint32_t acc_parity = 0;
  uint32_t p_phase = 0, c_phase = 0;           // p_phase: parity of the NEXT wait on empty[p_stage] (valid once produced >= N_STAGES)
  long long t_full = 0, t_empty = 0, t_acc = 0, t_epi = 0, t_in = 0, t_mma = 0, t_copy = 0, t_body = 0, t_all = clock64();   // diagnostic clocks (thread 0, dbg != nullptr)
  // thread 0: request weight tiles up to (but not including) index `upto`
  auto top_up = [&](int upto) {
    while (produced < upto && produced < total) {
      if (produced >= N_STAGES) {
        const long long c0 = dbg ? clock64() : 0;
        mbar_wait(&empty[p_stage], p_phase);
        if (dbg) t_empty += clock64() - c0;
      }
      const long long cc0 = dbg ? clock64() : 0;
      const BgymPolicyStep& ps = c_prog[p_prog];
      const uint32_t part = (uint32_t)ps.bytes / CTAS;            // a CTA of a pair holds half the rows of the weight tile
      if (elect_one()) {
        mbar_expect_tx(&full[p_stage], part);
        bulk_g2s(W + p_stage * STAGE_BYTES, weights + ps.offset + rank * part, part, &full[p_stage]);
      }
      __syncwarp();
      if (dbg) t_copy += clock64() - cc0;
      produced++;
      if (++p_prog == n_steps) p_prog = 0;
      if (++p_stage == N_STAGES) { p_stage = 0; if (produced > N_STAGES) p_phase ^= 1; }
    }
  };
  constexpr int PRODUCER_WARP = 1;            // warp 1 feeds the weight ring, warp 0 issues the MMAs (one elected lane each)
  if (warp == PRODUCER_WARP) top_up(N_STAGES - 1);

  for (long long t = 0; t < my_tiles; t++) {
    const long long tile = (cluster_id + t * n_clusters) * CTAS + rank;
    const long long c_in = dbg ? clock64() : 0;
    // ---- the tile's input, straight from the observation records (BalatroFeaturesExtractor.forward's preprocessing,
    // train_balatro_agent.py:84-113): four threads per env; thread (row, part) writes 14 of the 56 chunks of the 8-hot hand
    // block (K-blocks 0..6: column 52 slot + card = 1 for every occupied hand slot) and two chunks of K-block 7 (joker ids as
    // numbers at columns 0..9, the 21 scaled game scalars at columns 16..36).  bf16(1.0) = 0x3F80.
    {
      const int r = tid / PARTS, part = tid % PARTS;
      const long long g = tile * TILE_M + r;
      const uint8_t* o = obs + (g < n ? g : 0) * OBS_BYTES;
      const unsigned long long hand = g < n ? __ldg(reinterpret_cast<const unsigned long long*>(o)) : ~0ull;   // 8 x int8, -1 = empty
#pragma unroll
      for (int c = 0; c < 56 / PARTS; c++) *reinterpret_cast<uint4*>(X + x_offset(r, part * (56 / PARTS) + c)) = make_uint4(0, 0, 0, 0);
      __syncwarp();      // the four threads of a row are neighbours in one warp: zeros first, then the (at most eight) ones
#pragma unroll
      for (int k = 0; k < 8 / PARTS; k++) {
        const int slot = (8 / PARTS) * part + k;
        const int card = (int)(int8_t)(hand >> (8 * slot));
        if (card >= 0 && card < 52) {
          const int col = 52 * slot + card;
          *reinterpret_cast<uint16_t*>(X + x_offset(r, col >> 3) + 2 * (col & 7)) = (uint16_t)0x3F80;
        }
      }
      // K-block 7: chunks 56..63 of the row; part p writes chunks 56 + 2p, 57 + 2p
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; k++) v[k] = 0.0f;
      if (g < n) {
        auto i8f = [&](int off) { return (float)*reinterpret_cast<const int8_t*>(o + off); };
        auto i16f = [&](int off) { return (float)*reinterpret_cast<const int16_t*>(o + off); };
        const float r10 = 1.0f / 10.0f;
        const int pp = part * 4 / PARTS;      // which 16 columns of K-block 7 this thread's values belong to
        if (pp == 0) {              // columns 0..15: joker_ids[0..9]
#pragma unroll
          for (int k = 0; k < 10; k++) v[k] = i16f(64 + 2 * k);
        } else if (pp == 1) {       // columns 16..31: chips_scored/1e6, chips_needed/1e5, progress_ratio, money/100, ante/10, round/3,
                                    //                 hands_left/10, discards_left/5, hand_levels[0..7]/10
          v[0] = (float)*reinterpret_cast<const long long*>(o + 24) * (1.0f / 1e6f);
          v[1] = (float)*reinterpret_cast<const int*>(o + 44) * (1.0f / 1e5f);
          v[2] = *reinterpret_cast<const float*>(o + 36);
          v[3] = (float)*reinterpret_cast<const int*>(o + 48) * (1.0f / 100.0f);
          v[4] = i16f(60) * r10;
          v[5] = i8f(148) * (1.0f / 3.0f);
          v[6] = i8f(149) * r10;
          v[7] = i8f(150) * (1.0f / 5.0f);
#pragma unroll
          for (int k = 0; k < 8; k++) v[8 + k] = i8f(134 + k) * r10;
        } else if (pp == 2) {       // columns 32..47: hand_levels[8..11]/10, phase/3
#pragma unroll
          for (int k = 0; k < 4; k++) v[k] = i8f(142 + k) * r10;
          v[4] = i8f(155) * (1.0f / 3.0f);
        }
      }
      uint4 q0, q1;
      q0.x = pack_bf16x2(v[0], v[1]); q0.y = pack_bf16x2(v[2], v[3]); q0.z = pack_bf16x2(v[4], v[5]); q0.w = pack_bf16x2(v[6], v[7]);
      q1.x = pack_bf16x2(v[8], v[9]); q1.y = pack_bf16x2(v[10], v[11]); q1.z = pack_bf16x2(v[12], v[13]); q1.w = pack_bf16x2(v[14], v[15]);
      if (PARTS == 4) {
        *reinterpret_cast<uint4*>(X + x_offset(r, 56 + 2 * part)) = q0;
        *reinterpret_cast<uint4*>(X + x_offset(r, 57 + 2 * part)) = q1;
      } else {
        *reinterpret_cast<uint4*>(X + x_offset(r, 56 + part)) = (part & 1) ? q1 : q0;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    cta_sync();
    if (dbg) t_in += clock64() - c_in;
    int s = 0;
    for (int group = 0; group < N_GROUPS; group++) {
      // the group's steps: [s, s_end)
      int s_end = s;
      while (!c_prog[s_end].last) s_end++;
      s_end++;
      if (warp == PRODUCER_WARP) {
        // keep the ring N_STAGES - 1 tiles ahead of the LAST step of this group: every stage it waits for is freed by an MMA of
        // this group or an earlier one, so it never waits on work that needs the coming epilogue (which needs this thread)
        top_up((int)(t * n_steps) + s_end + N_STAGES - 1);
      } else if (warp == 0 && rank == 0) {
        tc_fence_after();
        for (; s < s_end; s++) {
          const BgymPolicyStep& ps = c_prog[s];
          const long long cf = dbg ? clock64() : 0;
          mbar_wait(&full[c_stage], c_phase);
          if (CTAS == 2) mbar_wait(&peer_full[c_stage], c_phase);
          if (dbg) t_full += clock64() - cf;
          tc_fence_after();
          const uint64_t a_desc = umma_desc(x_base + (uint32_t)ps.a_kb * KB_BYTES);
          const uint64_t b_desc = umma_desc(w_base + (uint32_t)c_stage * STAGE_BYTES);
          const uint32_t idesc = umma_idesc(ps.n, TILE_M * CTAS);
          const uint32_t d_addr = tmem_base + (uint32_t)ps.col;
          const long long cm = dbg ? clock64() : 0;
          if (elect_one()) {
            if (CTAS == 2) {
              umma_pair(d_addr, a_desc, b_desc, idesc, ps.first ? 0u : 1u);
              umma_pair(d_addr, a_desc + 2, b_desc + 2, idesc, 1u);
              umma_pair(d_addr, a_desc + 4, b_desc + 4, idesc, 1u);
              umma_pair(d_addr, a_desc + 6, b_desc + 6, idesc, 1u);
              tc_commit_pair(&empty[c_stage]);
            } else {
              umma(d_addr, a_desc, b_desc, idesc, ps.first ? 0u : 1u);      // four K = 16 slices of the 64-column block:
              umma(d_addr, a_desc + 2, b_desc + 2, idesc, 1u);              // + 32 bytes = + 2 in the descriptor's address field
              umma(d_addr, a_desc + 4, b_desc + 4, idesc, 1u);
              umma(d_addr, a_desc + 6, b_desc + 6, idesc, 1u);
              tc_commit(&empty[c_stage]);
            }
          }
          __syncwarp();
          if (dbg) t_mma += clock64() - cm;
          if (++c_stage == N_STAGES) { c_stage = 0; c_phase ^= 1; }
        }
        if (elect_one()) { if (CTAS == 2) tc_commit_pair(acc_bar); else tc_commit(acc_bar); }
        __syncwarp();
      } else if (CTAS == 2 && warp == 0) {
        // the peer's thread 0 tells the leader when this CTA's half of a weight tile has landed
        for (; s < s_end; s++) {
          mbar_wait(&full[c_stage], c_phase);
          if (elect_one()) mbar_arrive_remote(&peer_full[c_stage], 0);
          __syncwarp();
          if (++c_stage == N_STAGES) { c_stage = 0; c_phase ^= 1; }
        }
      }
      s = s_end;
      // Only warp 0 polls the accumulator barrier (its other lanes reconverge with thread 0 first); the other fifteen warps
      // sleep in the CTA barrier instead of spinning on mbarrier.try_wait next to the two issuing threads.
      __syncwarp();
      const long long ca = dbg ? clock64() : 0;
      if (warp == 0) mbar_wait(acc_bar, acc_parity);
      acc_parity ^= 1;
      __syncthreads();
      tc_fence_after();
      const long long ce = dbg ? clock64() : 0;
      t_acc += ce - ca;
      const float* gb = bias + bias_offset(group);
      if (group == 0) {          // first layers: 448 columns (hand 256 | joker 128 | game 64)
        epilogue_to_x<ACT_RELU>(taddr, row, slice_lo(448), slice_hi(448), gb, X);
      } else if (group == 1) {   // 224 columns + the 32 zero columns of combined_net.0's K padding
        epilogue_to_x<ACT_RELU>(taddr, row, slice_lo(224), slice_hi(224), gb, X);
        if (slice == SLICES - 1) {
#pragma unroll
          for (int q = 28; q < 32; q++) *reinterpret_cast<uint4*>(X + x_offset(row, q)) = make_uint4(0, 0, 0, 0);
        }
      } else if (group <= 3) {
        epilogue_to_x<ACT_RELU>(taddr, row, slice_lo(512), slice_hi(512), gb, X);
      } else if (group <= 5) {
        epilogue_to_x<ACT_TANH>(taddr, row, slice_lo(512), slice_hi(512), gb, X);
      } else if (slice == 0) {
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
      if (dbg) t_body += clock64() - ce;       // this warp's epilogue body, without the CTA barrier that follows
      fence_proxy_async();
      tc_fence_before();
      cta_sync();
      if (dbg) t_epi += clock64() - ce;
    }
  }
  if (dbg && tid == 0) {
    long long* d = dbg + blockIdx.x * 16;
    d[0] = clock64() - t_all; d[1] = t_in; d[2] = t_full; d[4] = t_acc; d[5] = t_epi; d[6] = my_tiles; d[7] = t_mma; d[9] = t_body;
  }
  if (dbg && tid == PRODUCER_WARP * 32) {
    long long* d = dbg + blockIdx.x * 16;
    d[3] = t_empty; d[8] = t_copy;
  }
  tc_fence_before();
  cta_sync();
  if (warp == 0) {
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
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

int bgym_policy_forward(const void* obs, const void* weights, const float* bias, float* logits, float* value, int64_t n, void* stream) {
  if (n < 0 || !obs || !weights || !bias || !logits || !value) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward: bad arguments"); return -1; }
  if (((uintptr_t)obs | (uintptr_t)weights | (uintptr_t)bias | (uintptr_t)logits) & 15) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward: obs / weights / bias / logits must be 16-byte aligned"); return -1; }
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
      if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_mlp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_mlp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
      if (e != cudaSuccess) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward setup: %s", cudaGetErrorString(e)); return (int)e; }
      g_ready[dev] = true;
    }
  }
  // BGYM_POLICY_CTAS=2: clusters of two CTAs, one cta_group::2 MMA (M = 256) per weight-tile slice; default: single CTAs
  static const int ctas = (getenv("BGYM_POLICY_CTAS") && atoi(getenv("BGYM_POLICY_CTAS")) == 2) ? 2 : 1;
  const long long tiles = (n + TILE_M - 1) / TILE_M;
  const long long groups = (tiles + ctas - 1) / ctas;
  const long long max_clusters = g_sms[dev] / ctas;
  const int grid = (int)((groups < max_clusters ? groups : max_clusters) * ctas);
  // BGYM_POLICY_CLOCK=1 (diagnostic): per-CTA clocks of thread 0 — input load, waits on weight tiles, on freed stages, on the
  // accumulators, and the epilogues — averaged over the CTAs and printed after a synchronisation
  static const bool clocks = getenv("BGYM_POLICY_CLOCK") != nullptr;
  long long* dbg = nullptr;
  if (clocks) { cudaMalloc(&dbg, (size_t)grid * 16 * sizeof(long long)); cudaMemset(dbg, 0, (size_t)grid * 16 * sizeof(long long)); }
  if (ctas == 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(N_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, policy_mlp_kernel<2>, reinterpret_cast<const uint8_t*>(obs), reinterpret_cast<const uint8_t*>(weights), bias,
                           logits, value, (long long)n, n_steps, dbg);
    if (e != cudaSuccess) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward launch (pairs): %s", cudaGetErrorString(e)); return (int)e; }
  } else {
    policy_mlp_kernel<1><<<grid, N_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(obs), reinterpret_cast<const uint8_t*>(weights),
                                                                              bias, logits, value, n, n_steps, dbg);
  }
  e = cudaGetLastError();
  if (clocks && e == cudaSuccess) {
    cudaStreamSynchronize((cudaStream_t)stream);
    long long* h = (long long*)malloc((size_t)grid * 16 * sizeof(long long));
    cudaMemcpy(h, dbg, (size_t)grid * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
    double sum[16] = {0};
    for (int b = 0; b < grid; b++) for (int k = 0; k < 10; k++) sum[k] += (double)h[b * 16 + k];
    const double tiles = sum[6] > 0 ? sum[6] : 1;
    fprintf(stderr, "[bgym policy clocks, cycles per tile] total %.0f | input %.0f | wait weights %.0f | wait stage %.0f | wait accumulators %.0f | epilogue %.0f | in the four MMA issues %.0f | copy issue (producer thread) %.0f | epilogue body of warp 0 %.0f\n",
            sum[0] / tiles, sum[1] / tiles, sum[2] / tiles, sum[3] / tiles, sum[4] / tiles, sum[5] / tiles, sum[7] / tiles, sum[8] / tiles, sum[9] / tiles);
    free(h); cudaFree(dbg);
  }
  if (e != cudaSuccess) { snprintf(g_perr, sizeof g_perr, "bgym_policy_forward launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

}  // extern "C"
