// Core of the next finder for patterns of the shape X* c Y* (C5: `\s*,\s*`) -- NOT part of the library yet: no
// product source includes this header.  It exists so that the arithmetic can be checked on the CPU
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

// 32 bits of a bit array starting at window position p (0 <= p, p + 32 <= 32 * kWords)
MRX_HD uint32_t take32(const uint32_t *w, int p) {
  const int k = p >> 5, s = p & 31;
  const uint32_t lo = w[k];
  const uint32_t hi = k + 1 < kWords ? w[k + 1] : 0u;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
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

// The separator that owns the comma at window position p (kPre <= p < kPre + kChunk).
// Returns 1 and (start, end) as window positions, 0 when the reference reports nothing for this comma;
// sets *overflow when a run is too long for the window.
MRX_HD int separator_at(const Window &w, int p, int *start, int *end, int *overflow) {
  const uint32_t below = take32(w.l, p - 32);                 // bit 31 = position p - 1
  const int back = clz32(~below);                             // leading ones
  const uint32_t above = take32(w.l, p + 1) & ~take32(w.bad, p + 1);  // Y = X = L \ bad; bit 0 = position p + 1
  const int fwd = ctz32(~above);                              // trailing ones
  if (back == 32 || fwd == 32) *overflow = 1;
  if (back > 0) {
    const uint32_t badbelow = take32(w.bad, p - 32);
    if (badbelow >> (32 - back)) return 0;                    // the attempt dies inside the back-scanned run
  }
  *start = p - back;
  *end = p + 1 + fwd;
  return 1;
}

}  // namespace sep
}  // namespace mrx
