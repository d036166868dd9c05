// Bit-parallel core of the fused split kernel (k_split in k_split.cu) for patterns of the shape X* c Y*
// (BASELINE config 5: `\s*,\s*`).  Host + device header; tests/test_sepbits_cpu.py compiles it with g++ into a
// harness that splits texts chunk by chunk exactly as the kernel's threads do and holds the pieces against the
// oracle.
//
// Reference semantics (HybridMatcher._match_all_required_byte, matcher.mojo:864-898, on the automaton
// S0 -X-> S1 -c-> S2(acc) -Y-> S3(acc), dfa.mojo:699-807), per occurrence p of c:
//   start = p - (length of the maximal run of L bytes ending at p - 1)      L = the back-scan set (first-class lookup)
//   no separator when that run holds a byte of L that X lacks ("bad": `\v` for `\s`)
//   end   = p + 1 + (length of the maximal run of Y bytes starting at p + 1)            Y = X (planner precondition)
// split (matcher.mojo:1357-1393) then keeps the bytes outside the separators; piece k + 1 starts where the k-th
// separator ends, so its offset in the output is the number of kept bytes before the k-th valid c.
//
// Everything is done on membership BITS of a window around one thread's 144-byte chunk: window bit i <-> text byte
// chunk_start - 96 + i, 9 words = 96 bits before the chunk, its 144, 48 after it.  Runs are followed with the
// carry of a multi-word addition (adding a bit at the low end of a run of ones clears the run and lands right
// after it):
//   valid c:      V = C & ~(L + bad)        a bad bit inside an L-run makes the sum land on the byte after the run
//   runs after:   A = Y & ~(Y + ((V << 1) & Y))
//   runs before:  the same in the bit-reversed window
//   kept        = ~(V | A | before);  a run that is both after one valid c and before the next one is the overlap
//                 on which the reference's split is undefined (matcher.mojo:1389).
// The window is wide enough whenever no run of L bytes is longer than 46; the kernel raises its overflow flag when
// two adjacent 16-byte groups (aligned to 16) are entirely in L -- true for every run of 47 or more -- and the host
// then takes the general path.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MRX_HD __host__ __device__ __forceinline__
#else
#define MRX_HD inline
#endif

namespace mrx {
namespace sepb {

constexpr int kWin = 9;        // window words
constexpr int kPreBits = 96;   // window bits before the chunk
constexpr int kOwn = 5;        // words that hold the chunk's 144 bits: window words 3..7 (the last one half)

// as it sits in HBM / shared memory: class bits of every byte value
struct Table {
  uint32_t T[256];  // bit 0: the byte is c;  bit 8: it is in L;  bit 16: it is in L but not in X
};

MRX_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
  return (x >> 16) | (x << 16);
#endif
}

struct Win { uint32_t w[kWin]; };

MRX_HD Win add(const Win &a, const Win &b) {
  Win r;
  uint32_t carry = 0;
#pragma unroll
  for (int k = 0; k < kWin; ++k) {
    const uint64_t s = (uint64_t)a.w[k] + b.w[k] + carry;
    r.w[k] = (uint32_t)s;
    carry = (uint32_t)(s >> 32);
  }
  return r;
}
MRX_HD Win shl1(const Win &a) {
  Win r;
#pragma unroll
  for (int k = 0; k < kWin; ++k) r.w[k] = (a.w[k] << 1) | (k ? a.w[k - 1] >> 31 : 0u);
  return r;
}
MRX_HD Win rev(const Win &a) {
  Win r;
#pragma unroll
  for (int k = 0; k < kWin; ++k) r.w[k] = brev32(a.w[kWin - 1 - k]);
  return r;
}
// runs of `y` that start right after a bit of `v`
MRX_HD Win runs_after(const Win &v, const Win &y) {
  const Win s1 = shl1(v);
  Win s, r;
#pragma unroll
  for (int k = 0; k < kWin; ++k) s.w[k] = s1.w[k] & y.w[k];
  const Win sum = add(y, s);
#pragma unroll
  for (int k = 0; k < kWin; ++k) r.w[k] = y.w[k] & ~sum.w[k];
  return r;
}

struct Own {
  uint32_t kept[kOwn];  // bytes of the chunk that belong to no separator (bits past the chunk / the text: caller masks)
  uint32_t sep[kOwn];   // the valid occurrences of c in the chunk
  uint32_t overlap;     // != 0: a run is shared by two separators
};

MRX_HD Own separators(const Win &C, const Win &L, const Win &Bad) {
  Win Y, V;
  const Win t = add(L, Bad);
#pragma unroll
  for (int k = 0; k < kWin; ++k) {
    Y.w[k] = L.w[k] & ~Bad.w[k];
    V.w[k] = C.w[k] & ~t.w[k];
  }
  const Win A = runs_after(V, Y);
  const Win Bf = rev(runs_after(rev(V), rev(Y)));
  Own o;
  o.overlap = 0;
#pragma unroll
  for (int j = 0; j < kOwn; ++j) {
    const int k = kPreBits / 32 + j;
    const uint32_t own = j == kOwn - 1 ? 0xffffu : 0xffffffffu;
    o.kept[j] = ~(V.w[k] | A.w[k] | Bf.w[k]) & own;
    o.sep[j] = V.w[k] & own;
    o.overlap |= A.w[k] & Bf.w[k] & own;
  }
  return o;
}

}  // namespace sepb
}  // namespace mrx
