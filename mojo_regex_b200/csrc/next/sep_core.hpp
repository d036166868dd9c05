// Core of the next finder for patterns of the shape X* c Y* (C5: `\s*,\s*`) -- NOT part of the library yet: no
// product source includes this header (only the experimental next/k_sep.cu does).  It exists so that the arithmetic can be checked on the CPU
// (tests/test_sep_core_cpu.py compiles it with g++ and holds it against the oracle) before it is wrapped in a
// kernel.  DESIGN.md section 8, item 1; the closed form it implements is in tests/test_finder_model_cpu.py.
//
// Everything works on membership BITS of a window around one thread's 144-byte chunk:
//   window position p = 0 .. 223  <->  text byte chunk_start - 32 + p   (7 words; the chunk is p = 32 .. 175)
//   C = the required byte, L = the back-scan set (first-class lookup, matcher.mojo:881-884), bad = L \ X,
//   Y = X (planner precondition).  Bits of bytes outside the text are 0.
// For a comma at p:  the 32 L-bits below p give the back-run length (count of leading ones), the 32 Y-bits above
// p the forward-run length (count of trailing ones); a run that fills its whole window is longer than the
// window can tell -> `overflow`, the caller falls back to the general kernel for this text.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MRX_HD __host__ __device__ __forceinline__
#else
#define MRX_HD inline
#endif

namespace mrx {
namespace sep {

constexpr int kChunk = 144;   // bytes per thread, as in the other kernels
constexpr int kPre = 32;      // window bits before the chunk
constexpr int kWords = 7;     // 224 window bits: 32 + 144 + 48

// how the sets are recognised on four packed bytes
struct Classes {
  uint32_t needle;        // c x 0x01010101
  uint32_t range_lo;      // L contains the bytes range_lo .. range_hi (both < 128); range_lo > range_hi: no range
  uint32_t range_hi;
  uint32_t extra[2];      // further single bytes of L, x 0x01010101
  int32_t n_extra;
  uint32_t bad[2];        // the bytes of L that X lacks, x 0x01010101
  int32_t n_bad;
};

// bit 7 of every byte of the result: that byte of w equals the splatted byte (exact, no cross-byte carries)
MRX_HD uint32_t eq_flags(uint32_t w, uint32_t splat) {
  const uint32_t x = w ^ splat;
  return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
// bit 7 of every byte: lo <= byte <= hi, for lo, hi < 128
MRX_HD uint32_t range_flags(uint32_t w, uint32_t lo, uint32_t hi) {
  const uint32_t t = w & 0x7f7f7f7fu;
  const uint32_t ge = t + (0x80u - lo) * 0x01010101u;       // bit 7 <=> low 7 bits >= lo
  const uint32_t gt = t + (0x7fu - hi) * 0x01010101u;       // bit 7 <=> low 7 bits >  hi
  return ge & ~gt & ~w & 0x80808080u;
}
// flags (bit 7 of each byte) -> 4 bits, byte 0 -> bit 0
MRX_HD uint32_t nibble(uint32_t flags) { return (((flags >> 7) * 0x01020408u) >> 24) & 0xfu; }

struct WordBits { uint32_t c, l, bad; };
MRX_HD WordBits classify4(uint32_t w, const Classes &k) {
  WordBits r;
  r.c = nibble(eq_flags(w, k.needle));
  uint32_t l = k.range_lo <= k.range_hi ? range_flags(w, k.range_lo, k.range_hi) : 0u;
  if (k.n_extra > 0) l |= eq_flags(w, k.extra[0]);
  if (k.n_extra > 1) l |= eq_flags(w, k.extra[1]);
  uint32_t b = 0;
  if (k.n_bad > 0) b |= eq_flags(w, k.bad[0]);
  if (k.n_bad > 1) b |= eq_flags(w, k.bad[1]);
  r.l = nibble(l);
  r.bad = nibble(b);
  return r;
}

struct Window { uint32_t c[kWords], l[kWords], bad[kWords]; };

// membership bits of 16 bytes given as four little-endian words (bit i = byte i); `live` = how many of the 16
// bytes exist in the text (bits of the others are cleared)
MRX_HD void bits16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int live, const Classes &k, uint32_t &c,
                   uint32_t &l, uint32_t &b) {
  const uint32_t ws[4] = {w0, w1, w2, w3};
  c = l = b = 0;
  for (int i = 0; i < 4; ++i) {
    const WordBits r = classify4(ws[i], k);
    c |= r.c << (4 * i);
    l |= r.l << (4 * i);
    b |= r.bad << (4 * i);
  }
  if (live < 16) {
    const uint32_t m = live <= 0 ? 0u : ((1u << live) - 1u);
    c &= m; l &= m; b &= m;
  }
}

// Bit streams of one tile (one uint16 per 16 text bytes, per set): halfwords [0, 2) = the 32 bytes before the
// tile, [2, 2 + 9 T) = the tile (T threads x 9 vectors), then 3 halfwords for the 48 bytes after it.
constexpr int kStreamPre = 2, kStreamPost = 3;
MRX_HD int stream_chunk(int tid, int vec) { return kStreamPre + 9 * tid + vec; }
MRX_HD int stream_post(int n_threads, int vec) { return kStreamPre + 9 * n_threads + vec; }
// the 224 window bits of thread tid: 14 halfwords starting 2 halfwords (32 bits) before its chunk
MRX_HD void window_from_streams(const uint16_t *bc, const uint16_t *bl, const uint16_t *bb, int tid, Window &w) {
  for (int k = 0; k < kWords; ++k) {
    const int h = 9 * tid + 2 * k;
    w.c[k] = (uint32_t)bc[h] | ((uint32_t)bc[h + 1] << 16);
    w.l[k] = (uint32_t)bl[h] | ((uint32_t)bl[h + 1] << 16);
    w.bad[k] = (uint32_t)bb[h] | ((uint32_t)bb[h + 1] << 16);
  }
}

MRX_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
MRX_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return x ? __ffs((int)x) - 1 : 32;
#else
  return x ? __builtin_ctz(x) : 32;
#endif
}

// The separator that owns the comma at bit j of window word k (window position p = 32 k + j), given the words
// k - 1, k, k + 1 of the L and bad streams -- the caller indexes them statically, so a kernel keeps the
// window in registers.  Returns 1 and the run lengths (start = p - back, end = p + 1 + fwd), 0 when the
// reference reports nothing for this comma; sets *overflow when a run fills its 32-bit window (longer than
// the window can tell).
MRX_HD int separator_at(uint32_t l_prev, uint32_t l_cur, uint32_t l_next, uint32_t b_prev, uint32_t b_cur,
                        uint32_t b_next, int j, int *back, int *fwd, int *overflow) {
  // the 32 L-bits below p: bit 31 = position p - 1
  const uint32_t below = (uint32_t)(((((uint64_t)l_cur) << 32) | l_prev) >> j);
  const int nb = clz32(~below);  // leading ones
  // the 32 Y-bits above p (Y = X = L minus bad): bit 0 = position p + 1
  const uint64_t y2 = ((((uint64_t)(l_next & ~b_next)) << 32) | (l_cur & ~b_cur));
  const uint32_t above = (uint32_t)(y2 >> (j + 1));
  const int nf = ctz32(~above);  // trailing ones
  if (nb == 32 || nf == 32) *overflow = 1;
  *back = nb;
  *fwd = nf;
  if (nb > 0) {
    const uint32_t badbelow = (uint32_t)(((((uint64_t)b_cur) << 32) | b_prev) >> j);
    if (badbelow >> (32 - nb)) return 0;  // the attempt dies inside the back-scanned run
  }
  return 1;
}

}  // namespace sep
}  // namespace mrx
