// Device-side building blocks shared by the kernels (sm_100a).
//
//  * DevProg: the packed automaton as it sits in HBM and, per CTA, in shared memory.
//  * Tile staging with the TMA 1-D bulk copy (cp.async.bulk + mbarrier; SASS: UBLKCP).
//  * SWAR byte filters (4 bytes per 32-bit ALU op) producing 0x80-per-byte flag words.
//  * Decoupled look-back over 64-bit tile descriptors (single-pass ordered compaction).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mrx {

constexpr int kThreads = 256;
constexpr int kChunk = 144;                    // bytes per thread: 9 x 16 B; 9 is odd, so the 8 lanes of a
                                               // quarter-warp hit 8 distinct 16-B bank groups on LDS.128
constexpr int kVecs = kChunk / 16;             // 9
constexpr int kTile = kThreads * kChunk;       // 36864 bytes of text per CTA
constexpr int kHalo = 256;                     // look-ahead bytes staged after the tile
constexpr int kMaxTrans = 8192;                // n_states * n_classes budget in shared memory
constexpr int kMaxLiteral = 256;

enum DevFlags : int32_t {
  DF_END_ANCHOR = 1 << 0,
  DF_START_ANCHOR = 1 << 1,
  DF_START_ACCEPTING = 1 << 2,
  DF_SHORTCUT = 1 << 3,  // _try_match_simd is live inside attempts (dfa.mojo:1956-1968)
  DF_DENSE_FWD = 1 << 4, // `trans` holds the dense forward table (uint16 row offsets, see RevTable::fwd)
                         // instead of the class-compressed one; `fwd_start` = row of the start state
};

struct DevProg {
  int32_t n_states, n_classes, start_state, flags;
  int32_t literal_len, required_byte, trans_len, fwd_start;
  uint8_t byte_class[256];
  uint8_t lookup[256];    // CharacterClassSIMD.lookup_table of the engine's matcher
  uint8_t cand[256];      // candidate predicate of the active plan
  uint8_t accepting[256];
  uint8_t literal[kMaxLiteral];
  uint8_t trans[kMaxTrans];
};

// ---- PTX helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_span(int64_t *out, int64_t s, int64_t e) {
  asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(out), "l"(s), "l"(e) : "memory");
}

// ---- SWAR filters: input word of 4 text bytes -> 0x80 in every byte that passes --------------
constexpr uint32_t kH = 0x80808080u, kL = 0x01010101u;

// exact per-byte equality with a splatted needle
__device__ __forceinline__ uint32_t swar_eq(uint32_t x, uint32_t splat) {
  uint32_t t = x ^ splat;                       // zero byte <=> equal
  uint32_t u = ((t & ~kH) + ~kH) | t;           // bit7 set <=> byte non-zero (no cross-byte carry)
  return ~u & kH;
}
// lo <= byte <= hi for lo, hi < 0x80 handled exactly; bytes >= 0x80 handled via the general form
__device__ __forceinline__ uint32_t swar_ge(uint32_t x, uint32_t c_splat) {
  uint32_t z = (x | kH) - (c_splat & ~kH);      // bit7 <=> (x & 0x7f) >= (c & 0x7f), per byte
  return ((x & ~c_splat) | (~(x ^ c_splat) & z)) & kH;
}
__device__ __forceinline__ uint32_t swar_range(uint32_t x, uint32_t lo_splat, uint32_t hi1_splat, bool hi_is_ff) {
  uint32_t ge = swar_ge(x, lo_splat);
  if (hi_is_ff) return ge;
  return ge & ~swar_ge(x, hi1_splat);
}
// flag word (0x80 per byte) -> 4-bit mask, bit i = byte i
__device__ __forceinline__ uint32_t flags_to_nibble(uint32_t f) { return (((f >> 7) * 0x00204081u) >> 21) & 15u; }

struct Filter {
  int32_t kind;  // FilterKind
  uint32_t n0, n1, n2, n3;  // splatted needles, or lo / hi+1 for the range form
  int32_t n_needles;
  int32_t hi_is_ff;
};

// ---- decoupled look-back --------------------------------------------------------------------
// desc[tile] = (flag << 62) | value; flag 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix.
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPre = 2ull << 62, kValMask = (1ull << 62) - 1;

// Called by all 32 lanes of one warp.  Publishes this tile's aggregate, walks predecessors
// 32 at a time, publishes the inclusive prefix, returns the exclusive prefix.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long *desc, int tile,
                                                                 unsigned long long aggregate) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(desc, kFlagPre | aggregate);
    return 0;
  }
  if (lane == 0) st_volatile_u64(desc + tile, kFlagAgg | aggregate);
  unsigned long long excl = 0;
  int idx = tile - 1;
  while (true) {
    const int my = idx - lane;
    unsigned long long d = kFlagPre;  // lanes past the beginning behave like a zero prefix
    if (my >= 0) {
      do { d = ld_volatile_u64(desc + my); } while ((d >> 62) == 0);
    }
    const unsigned pre_mask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
    const int first_pre = pre_mask ? (__ffs(pre_mask) - 1) : 32;
    unsigned long long v = (lane <= first_pre) ? (d & kValMask) : 0;  // aggregates up to and incl. the prefix lane
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (pre_mask) break;
    idx -= 32;
  }
  if (lane == 0) st_volatile_u64(desc + tile, kFlagPre | (excl + aggregate));
  return excl;
}

// The same look-back split in two, for kernels that publish a tile's aggregate as soon as it is known
// and resolve its prefix later (one tile behind), so that nobody waits on a predecessor.
__device__ __forceinline__ void lookback_publish(unsigned long long *desc, int tile, unsigned long long aggregate) {
  st_volatile_u64(desc + tile, (tile == 0 ? kFlagPre : kFlagAgg) | aggregate);
}
// Block-wide resolve (all kThreads threads): looks at 256 predecessors per round, so a whole wave of
// concurrently running tiles is crossed in two or three rounds.  Needs every predecessor's AGGREGATE to
// be published (true one tile behind); never waits for a predecessor's prefix.  `ws`: 2 * kThreads/32 + 2
// unsigned long long of shared memory.  Returns the exclusive prefix to every thread.
__device__ __forceinline__ unsigned long long lookback_resolve_block(unsigned long long *desc, int tile,
                                                                     unsigned long long aggregate,
                                                                     unsigned long long *ws) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int W = kThreads / 32;
  unsigned long long excl = 0;
  int idx = tile - 1;
  while (idx >= 0) {
    const int my = idx - tid;
    unsigned long long d = kFlagPre;  // positions before tile 0 act as a zero prefix
    if (my >= 0) {
      do { d = ld_volatile_u64(desc + my); } while ((d >> 62) == 0);
    }
    const unsigned pre_mask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
    const int first_pre = pre_mask ? (__ffs(pre_mask) - 1) : 32;
    unsigned long long v = (lane <= first_pre) ? (d & kValMask) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) { ws[warp] = v; ws[W + warp] = pre_mask ? 1ull : 0ull; }
    __syncthreads();
    bool found = false;
    for (int w = 0; w < W; ++w) {  // warps are ordered nearest-first
      excl += ws[w];
      if (ws[W + w]) { found = true; break; }
    }
    __syncthreads();
    if (found) break;
    idx -= kThreads;
  }
  if (tid == 0 && tile > 0) st_volatile_u64(desc + tile, kFlagPre | (excl + aggregate));
  return excl;
}

// Block-wide exclusive scan of one int per thread (kThreads threads). Returns exclusive value;
// *total gets the block sum.  `ws` needs kThreads/32 + 1 ints of shared memory.
__device__ __forceinline__ int block_exclusive_scan(int v, int *ws, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kThreads / 32 ? ws[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < kThreads / 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    if (lane < kThreads / 32) ws[lane] = winc - w;
    if (lane == kThreads / 32 - 1) ws[kThreads / 32] = winc;
  }
  __syncthreads();
  const int res = ws[warp] + inc - v;
  *total = ws[kThreads / 32];
  __syncthreads();
  return res;
}

}  // namespace mrx
