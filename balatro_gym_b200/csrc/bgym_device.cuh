// bgym_device.cuh — device-side building blocks shared by the sm_100a kernels:
// PTX wrappers (mbarrier, 1-D bulk async copies = TMA without a tensor map), Philox4x32-10,
// packed-byte helpers, constant tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bgym.h"
#include "../../include/bgym_tables.h"

namespace bgym {

// ---------------------------------------------------------------------------------------------
// constant tables (values generated from the reference's data tables, include/bgym_tables.h)
// ---------------------------------------------------------------------------------------------
__constant__ uint8_t c_joker_cost[BGYM_NUM_JOKERS + 1] = BGYM_JOKER_COST_INIT;
__constant__ BgymJokerFx c_joker_fx[BGYM_NUM_JOKERS + 1] = BGYM_JOKER_FX_INIT;
__constant__ int c_base_chips[12] = BGYM_BASE_CHIPS_INIT;
__constant__ int c_base_mult[12] = BGYM_BASE_MULT_INIT;
__constant__ int c_blind_chips[8][3] = BGYM_BLIND_CHIPS_INIT;
__constant__ int c_pack_cost[5] = BGYM_PACK_COST_INIT;
__constant__ int c_voucher_cost[2] = BGYM_VOUCHER_COST_INIT;
__constant__ double c_pow_1_15[101] = BGYM_POW_1_15_INIT;
__constant__ double c_pow_0_8[9] = BGYM_POW_0_8_INIT;
__constant__ double c_pow_1_5[101] = BGYM_POW_1_5_INIT;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk async copy (SASS: UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk async copy
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16-byte asynchronous global -> shared copy (SASS: LDGSTS), L2 only
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void sts128(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
// record accesses with a cache operator.  CG = the address is GLOBAL memory and the data is touched once by this SM
// (a list tile's hot / toggle / observation records): loads come straight from L2 and stores go there without
// taking an L1 line each, so that L1 keeps what a tile re-reads (its cold records, its stack frames)
template <bool CG>
__device__ __forceinline__ uint4 ldr128(const void* p) {
  if (CG) return __ldcg(reinterpret_cast<const uint4*>(p));
  return *reinterpret_cast<const uint4*>(p);
}
template <bool CG>
__device__ __forceinline__ void str128(void* p, uint4 v) {
  if (CG) __stcg(reinterpret_cast<uint4*>(p), v);
  else *reinterpret_cast<uint4*>(p) = v;
}
template <bool CG>
__device__ __forceinline__ void str64(void* p, uint2 v) {
  if (CG) __stcg(reinterpret_cast<uint2*>(p), v);
  else *reinterpret_cast<uint2*>(p) = v;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based; Salmon et al. SC'11 constants)
// ---------------------------------------------------------------------------------------------
#define BGYM_PHILOX_KEY1 0xB200CAFEu
#define BGYM_POLICY_KEY1 0x5A17AC71u
#define BGYM_SHUFFLE_KEY1 0xB200DECCu
#define BGYM_SAMPLE_KEY1 0xCA7E6031u

// inline form, for code that generates several independent blocks back to back (the reset's shuffle and modifier
// draws): two calls side by side give the scheduler two dependency chains to interleave
__device__ __forceinline__ uint4 philox4x32_10_inl(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// noinline on purpose: draws are rare (a few percent of env-steps) and the step kernel has ~20 draw
// sites; one shared copy keeps the kernel inside the instruction cache.
__device__ __noinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
  // fully unrolled: the rolled loop spent 5 of its 11 instructions per round on moves, the trip count and the
  // branch; the key schedule folds into constants
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Draw source of one env for one step: native Philox word stream (seed/counter in the state) or
// replay of the reference's recorded draws (BgymDraws).  Left-over words of the current block are
// dropped at the end of the call; the counter counts blocks.
struct Draws {
  const BgymDraws* tape;  // nullptr = native
  uint32_t seed, ctr;
  uint32_t buf0, buf1, buf2, buf3;
  uint32_t first;
  int pos, iu, ik;

  __device__ __forceinline__ void init(uint32_t seed_, uint32_t ctr_, const BgymDraws* tape_) {
    tape = tape_; seed = seed_; ctr = ctr_; first = ctr_; pos = 4; iu = 0; ik = 0;
    buf0 = buf1 = buf2 = buf3 = 0;
  }
  // Generate the first block now, with the whole warp converged, instead of inside whichever divergent
  // branch draws first.  The stream is unchanged: an unused prefetched block is not counted (blocks()).
  __device__ __forceinline__ void prefetch() {
    uint4 b = philox4x32_10(ctr, 0, 0, 0, seed, BGYM_PHILOX_KEY1);
    if (!tape) { ctr++; buf0 = b.x; buf1 = b.y; buf2 = b.z; buf3 = b.w; pos = 0; }
  }
  // number of blocks consumed so far (= the counter to store back)
  __device__ __forceinline__ uint32_t blocks() const { return (pos == 0 && ctr == first + 1) ? first : ctr; }
  __device__ __forceinline__ uint32_t word() {
    if (pos == 4) {
      uint4 b = philox4x32_10(ctr++, 0, 0, 0, seed, BGYM_PHILOX_KEY1);
      buf0 = b.x; buf1 = b.y; buf2 = b.z; buf3 = b.w; pos = 0;
    }
    uint32_t w = pos == 0 ? buf0 : pos == 1 ? buf1 : pos == 2 ? buf2 : buf3;
    pos++;
    return w;
  }
  __device__ double u01();
  __device__ int below(int n);
  __device__ int count_u01_below(int n, uint32_t tested, double threshold);
  // A step whose round advance is deferred to a second-level tile (bgym_step_part.cuh) continues the SAME draw
  // sequence there.  16 bits say where it stands: words used of the current block (0..4) | tape positions |
  // "the prefetched block is still untouched" (see blocks()); the block counter travels in the hot record.
  __device__ __forceinline__ uint32_t snapshot() const {
    return (uint32_t)pos | ((uint32_t)iu << 3) | ((uint32_t)ik << 8) | ((pos == 0 && ctr == first + 1) ? 0x4000u : 0u);
  }
  __device__ __forceinline__ void restore(uint32_t seed_, uint32_t ctr_, const BgymDraws* tape_, uint32_t snap) {
    tape = tape_; seed = seed_; ctr = ctr_; pos = (int)(snap & 7u); iu = (int)((snap >> 3) & 31u); ik = (int)((snap >> 8) & 63u);
    first = (snap & 0x4000u) ? ctr_ - 1 : ctr_;
    buf0 = buf1 = buf2 = buf3 = 0;
    if (!tape && pos < 4) {          // a block is open: regenerate it
      uint4 b = philox4x32_10(ctr_ - 1, 0, 0, 0, seed, BGYM_PHILOX_KEY1);
      buf0 = b.x; buf1 = b.y; buf2 = b.z; buf3 = b.w;
    }
  }
};

// Out of line (one copy each): the step kernel has ~25 draw sites, all on rare paths.
// CPython random.random(): (a >> 5, b >> 6) -> 53 bits
__device__ __noinline__ double Draws::u01() {
  if (tape) return tape->u[iu++];
  uint32_t a = word() >> 5;
  uint32_t b = word() >> 6;
  return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}
// the next n u01() draws in one call: how many of those marked in `tested` (bit j = draw j) fall below `threshold`.
// Same stream positions as n calls of u01(); the lanes of a tile that have draws to make run this loop together.
__device__ __noinline__ int Draws::count_u01_below(int n, uint32_t tested, double threshold) {
  int cnt = 0;
  if (tape) {
#pragma unroll 1
    for (int j = 0; j < n; j++) cnt += ((tested >> j) & 1u) && tape->u[iu + j] < threshold;
    iu += n;
    return cnt;
  }
#pragma unroll 1
  for (int j = 0; j < n; j++) {
    const uint32_t a = word() >> 5, b = word() >> 6;
    if ((tested >> j) & 1u) cnt += (a * 67108864.0 + b) * (1.0 / 9007199254740992.0) < threshold;
  }
  return cnt;
}
// unbiased integer in [0, n): Lemire multiply-shift with rejection
__device__ __noinline__ int Draws::below(int n) {
  if (tape) return tape->k[ik++];
  uint32_t un = (uint32_t)n;
  uint64_t m = (uint64_t)word() * un;
  uint32_t l = (uint32_t)m;
  if (l < un) {
    uint32_t t = (0u - un) % un;
    while (l < t) {
      m = (uint64_t)word() * un;
      l = (uint32_t)m;
    }
  }
  return (int)(m >> 32);
}

// j for Fisher-Yates position i (1..51) of the native shuffle: Philox block (i-1)/2 keyed
// (seed, shuffle key), words (x,y) for odd i, (z,w) for even i; Lemire multiply-shift, second word
// on rejection.  Independent per i, so a warp computes all 51 in parallel.
__device__ __forceinline__ int shuffle_j_from_block(uint4 b, int i) {
  uint32_t w0 = ((i - 1) & 1) ? b.z : b.x, w1 = ((i - 1) & 1) ? b.w : b.y;
  uint32_t un = (uint32_t)(i + 1);
  uint64_t m = (uint64_t)w0 * un;
  uint32_t l = (uint32_t)m;
  if (l < un) {                       // p < 52 / 2^32: only then is the exact threshold needed
    if (l < (0u - un) % un) m = (uint64_t)w1 * un;
  }
  return (int)(m >> 32);
}

// x / y for y > 0 and x >= 0 where x is often exactly zero (progress of a round that has just begun, a score of 0):
// the fp64 division's fast path does not take a zero numerator — every such lane went through the ~60-instruction
// out-of-line path, one small group of lanes at a time (ncu: 5.6 % of the PLAY list kernel's instructions)
// (a branch around the division is folded back into an unconditional one by the compiler, slow path included, and so is
// a plain select: the zero numerator is replaced by the denominator behind an empty asm — y / y takes the fast path —
// and the quotient by zero)
__device__ __forceinline__ double div_nz(double x, double y) {
  const bool nz = x != 0.0;
  double xs = nz ? x : y;
  asm volatile("" : "+d"(xs));      // opaque to the optimiser, which otherwise proves the select away and divides x again
  const double q = xs / y;
  return nz ? q : 0.0;
}

// ---------------------------------------------------------------------------------------------
// packed-byte helpers (8 x u8 in a u64, 8 x u4 in a u32)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int byte_at(uint64_t v, int i) { return (int)((v >> (8 * i)) & 0xFF); }
__device__ __forceinline__ uint64_t with_byte(uint64_t v, int i, int b) {
  return (v & ~(0xFFull << (8 * i))) | ((uint64_t)(b & 0xFF) << (8 * i));
}
__device__ __forceinline__ int nib_at(uint32_t v, int i) { return (int)((v >> (4 * i)) & 15); }
// the ordered selection list: n hand slots (0..7, each at most once) as the low n nibbles of a word.  Both helpers are
// branch-free: as loops over the list they ran to the longest list of the warp and were a quarter of the main pass's
// instructions (ncu).
// bit set of the listed slots
__device__ __forceinline__ uint32_t sel_slot_mask(uint32_t order, int n) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) m |= (k < n ? 1u : 0u) << ((order >> (4 * k)) & 15u);
  return m;
}
// position of `slot` in the list, -1 when it is not there (zero-nibble test on order ^ slot-in-every-nibble; the
// lowest flagged nibble is always a true match, borrows can only flag nibbles above it)
__device__ __forceinline__ int sel_find(uint32_t order, int n, int slot) {
  const uint32_t x = order ^ (0x11111111u * (uint32_t)slot);
  uint32_t t = (x - 0x11111111u) & ~x & 0x88888888u;
  t &= n >= 8 ? 0xFFFFFFFFu : ((1u << (4 * n)) - 1u);
  return t ? (__ffs((int)t) - 1) >> 2 : -1;
}
// position of the k-th (0-based) set bit of m, k < popc(m): a branch-free binary search on popcounts (as a loop that
// clears k low bits it ran to the largest k of the warp: the random policy's choice among ~10 legal actions)
__device__ __forceinline__ int select_bit64(uint64_t m, int k) {
  uint32_t w = (uint32_t)m;
  int pos = 0, c = __popc(w);
  if (k >= c) { k -= c; w = (uint32_t)(m >> 32); pos = 32; }
#pragma unroll
  for (int sh = 16; sh >= 1; sh >>= 1) {
    c = __popc(w & ((1u << sh) - 1u));
    if (k >= c) { k -= c; w >>= sh; pos += sh; }
  }
  return pos;
}
// 4 mask bits -> 4 bytes of 0/1
__device__ __forceinline__ uint32_t spread4(uint32_t bits) { return ((bits & 0xF) * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// card16 fields
__device__ __forceinline__ int c16_code(int c) { return c & 63; }
__device__ __forceinline__ int c16_enh(int c) { return (c >> 6) & 15; }
__device__ __forceinline__ int c16_edition(int c) { return (c >> 10) & 7; }
__device__ __forceinline__ int c16_seal(int c) { return (c >> 13) & 7; }

// chip value of one card: Rank.base_chips (cards.py:52-60) + CardState.calculate_chip_bonus
// (cards.py:262-267).  r = code >> 2 = rank - 2.
__device__ __forceinline__ int card_chips(int code, int enh, int edition) {
  int r = code >> 2;
  int v = min(r + 2, 10) + (r == 12);
  v += (enh == BGYM_ENH_BONUS) ? 30 : 0;
  v += (enh == BGYM_ENH_STONE) ? 50 : 0;
  v += (edition == BGYM_ED_FOIL) ? 50 : 0;
  return v;
}

// ---------------------------------------------------------------------------------------------
// poker-hand classification, balatro_game.py:40-93, on register histograms:
//   cnt  : 13 rank counters, 4 bits each (counts <= 4 because the cards are distinct)
//   rmask: rank presence bits, smask: suit presence bits, n: number of cards
// ---------------------------------------------------------------------------------------------
struct HandHist {
  uint64_t cnt;
  uint32_t rmask, smask;
  int n;
  __device__ __forceinline__ void clear() { cnt = 0; rmask = 0; smask = 0; n = 0; }
  __device__ __forceinline__ void add(int code) {
    int r = code >> 2;
    cnt += 1ull << (4 * r);
    rmask |= 1u << r;
    smask |= 1u << (code & 3);
    n++;
  }
};

__device__ __forceinline__ int classify(const HandHist& h) {
  if (h.n == 0) return BGYM_HT_HIGH_CARD;
  const uint64_t ones = 0x1111111111111ull;
  uint64_t b0 = h.cnt & ones, b1 = (h.cnt >> 1) & ones, b2 = (h.cnt >> 2) & ones;
  uint64_t is4 = b2, is3 = b0 & b1, is2 = b1 & ~b0;
  int n2 = __popcll(is2), n3 = __popcll(is3);
  bool flush = (__popc(h.smask) == 1) && h.n >= 5;
  uint32_t rm = h.rmask;
  bool straight = ((rm & (rm >> 1) & (rm >> 2) & (rm >> 3) & (rm >> 4)) != 0) || ((rm & 0x100Fu) == 0x100Fu);
  straight = straight && (__popc(rm) >= 5);
  if (straight && flush) return BGYM_HT_STRAIGHT_FLUSH;
  if (is4) return BGYM_HT_FOUR_KIND;
  if (n3 == 1 && n2 >= 1) return BGYM_HT_FULL_HOUSE;   // sorted counts start [3, 2, ...]
  if (flush) return BGYM_HT_FLUSH;
  if (straight) return BGYM_HT_STRAIGHT;
  if (n3 >= 1) return BGYM_HT_THREE_KIND;
  if (n2 >= 2) return BGYM_HT_TWO_PAIR;
  if (n2 == 1) return BGYM_HT_ONE_PAIR;
  return BGYM_HT_HIGH_CARD;
}

// RULES evaluator (BGYM_SCORE_RULES): BalatroSimulator.evaluate_hand balatro_sim.py:110-400 on register histograms.
//   cnt: 13 rank counters of 4 bits (up to 8 equal cards), scnt: 4 suit counters of 8 bits, n: cards
__device__ __forceinline__ int classify_rules(uint64_t cnt, uint32_t scnt, uint32_t rmask, int n, bool four_fingers, bool shortcut) {
  const uint64_t ones = 0x1111111111111ull;
  const uint64_t b0 = cnt & ones, b1 = (cnt >> 1) & ones, b2 = (cnt >> 2) & ones, b3 = (cnt >> 3) & ones;
  const uint64_t lo = ~b3 & ones;                                   // counts below 8
  const int n2 = __popcll(lo & ~b2 & b1 & ~b0), n3 = __popcll(lo & ~b2 & b1 & b0);   // exactly 2 / 3 (get_x_same :118)
  const bool n4 = (lo & b2 & ~b1 & ~b0) != 0, n5 = (lo & b2 & ~b1 & b0) != 0;       // exactly 4 / 5
  const int required = four_fingers ? 4 : 5;
  const bool sized = n <= 5 && n >= required;                       // :133, :155
  bool flush = false;
#pragma unroll
  for (int s = 0; s < 4; s++) flush |= (int)((scnt >> (8 * s)) & 0xFF) >= required;
  flush = flush && sized;
  bool straight = false;
  if (sized) {
    int length = 0;
    bool skipped = false;
#pragma unroll 1
    for (int r = 12; r >= 0 && !straight; r--) {                    // ranks 14 .. 2 (:173-187)
      if ((rmask >> r) & 1) length++;
      else if (shortcut && !skipped) skipped = true;
      else { length = 0; skipped = false; }
      straight = length >= required;
    }
    if (!straight) {                                                // wheel A-2-3-4-5, `skipped` carried over (:190-206)
      int wheel = 0;
      bool go = true;
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const int r = k == 0 ? 12 : k - 1;
        if (go) {
          if ((rmask >> r) & 1) wheel++;
          else if (shortcut && !skipped) skipped = true;
          else go = false;
        }
      }
      straight = wheel >= required;
    }
  }
  if (n5 && flush) return BGYM_HT_FLUSH_FIVE;
  if (n3 && n2 && flush) return BGYM_HT_FLUSH_HOUSE;
  if (n5) return BGYM_HT_FIVE_KIND;
  if (flush && straight) return BGYM_HT_STRAIGHT_FLUSH;
  if (n4) return BGYM_HT_FOUR_KIND;
  if (n3 && n2) return BGYM_HT_FULL_HOUSE;
  if (flush) return BGYM_HT_FLUSH;
  if (straight) return BGYM_HT_STRAIGHT;
  if (n3) return BGYM_HT_THREE_KIND;
  if (n2 == 2 || (n3 == 1 && n2 == 1)) return BGYM_HT_TWO_PAIR;
  if (n2) return BGYM_HT_ONE_PAIR;
  return BGYM_HT_HIGH_CARD;
}

// ScoreEngine.get_hand_chips_mult scoring_engine.py:87-101 (engine level = min(level, 15))
__device__ __forceinline__ void hand_base(int ht, int level, int& chips, int& mult) {
  int lv = min(max(level, 1), 15) - 1;
  chips = c_base_chips[ht] + 10 * lv;
  mult = c_base_mult[ht] + lv;
}

}  // namespace bgym
