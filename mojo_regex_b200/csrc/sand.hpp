// Shift-And formulation of the literal-set scan (k_sand in kernels.cu): the bit-parallel automaton of a small,
// prefix-free, nesting-free literal set (BASELINE config 2: hello | world | foo).  Host + device header; the
// table builder and the per-thread scan are checked on the CPU by tests/test_sand_core_cpu.py (g++ harness
// that walks a text exactly the way one kernel thread does) against the oracle.
//
// K slots (K = 1, 2 or 4 >= number of literals).  Literal i keeps its LAST state in bit 31 - i and the state j
// characters before its end in bit 31 - i - K*j.  One step per text byte:   D = ((D << K) + INIT) & B[byte]
//   * the shift moves every literal's states one character forward; what leaves bit 31 are states after the
//     last one;
//   * INIT holds the first-state bit of every literal.  No shifted bit can land on one (same residue mod K =
//     same literal, and nothing precedes a first state), so OR and ADD coincide: one IMAD;
//   * "literal i ends at this byte" <=> bit 31 - i of D: the top K bits of D.  They are funnel-shifted into an
//     accumulator (one SHF per byte), so after 32 / K bytes an accumulator word holds the hit bits of those
//     bytes, first byte on top: bit 31 - K*b - i <=> literal i ends at byte b.
// Tail form (Table::tail, used whenever K * (max_len + 3) <= 32): every literal's chain is extended by three
// "tail" states that accept any byte, so a hit stays visible in D for three more steps and moves up by K bits per
// step.  After the four bytes of a text word the top 4K bits of D are [hits of byte 0 | byte 1 | byte 2 | byte 3]
// -- the same layout four pushes would have produced -- and four bytes later they have left the register: one
// funnel shift per WORD instead of one per byte, and no clearing.
// The lookup address depends on the text only -- no dependent chain of shared-memory loads -- and B has 4-byte
// entries, so bytes that differ in their low five bits (all of a..z and the blank) sit in different banks: one
// wavefront per lookup on ordinary text, against 2.04 for the state-indexed table of k_dense
// (profiles/r01_k_dense_h_ncu.txt).
// Preconditions (program.cpp: build_dense enforces them): no literal is a proper prefix of another or occurs
// inside another, so at most one literal ends at any byte.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#if defined(__CUDACC__)
#define MRX_HD __host__ __device__ __forceinline__
#else
#define MRX_HD inline
#endif

namespace mrx {
namespace sand {

// as it sits in HBM / shared memory
struct Table {
  int32_t k = 0;        // slots: 1, 2 or 4
  int32_t n_lits = 0;
  int32_t max_len = 0;
  int32_t warm_words = 0;   // ceil((max_len - 1) / 4): 4-byte words of history that make a state exact
  uint32_t init = 0;
  uint32_t lens = 0;        // literal length by slot, 8 bits each
  int32_t tail = 0;         // 1: tail form (hit bits are read off D once per text word)
  uint32_t pad = 0;
  uint32_t B[256] = {0};
};

inline bool build(const std::vector<std::string> &lits, Table &t, bool allow_tail = true) {
  t = Table();
  const int n = (int)lits.size();
  if (n < 1 || n > 4) return false;
  const int k = n == 1 ? 1 : n == 2 ? 2 : 4;
  size_t max_len = 0;
  for (const std::string &w : lits) {
    if (w.empty() || w.size() > 255) return false;
    max_len = w.size() > max_len ? w.size() : max_len;
  }
  if ((size_t)k * max_len > 32) return false;
  t.k = k; t.n_lits = n; t.max_len = (int)max_len;
  t.warm_words = (int)((max_len - 1 + 3) / 4);
  t.tail = (allow_tail && (size_t)k * (max_len + 3) <= 32) ? 1 : 0;
  const int lift = t.tail ? 3 : 0;  // states sit 3 steps lower: the top 3K bits are the tails
  for (int i = 0; i < n; ++i) {
    const std::string &w = lits[(size_t)i];
    const int len = (int)w.size();
    for (int j = 0; j < len; ++j) t.B[(unsigned char)w[(size_t)(len - 1 - j)]] |= 1u << (31 - i - k * (j + lift));
    t.init |= 1u << (31 - i - k * (len - 1 + lift));
    t.lens |= (uint32_t)len << (8 * i);
  }
  if (t.tail)
    for (int c = 0; c < 256; ++c)
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) t.B[c] |= 1u << (31 - i - k * d);
  return true;
}

template <int K>
MRX_HD uint32_t step(uint32_t d, uint32_t b_entry, uint32_t init) {
  return ((d << K) + init) & b_entry;
}
// acc' = (acc << BITS) | top BITS bits of d   (BITS = K: one byte's hit bits; BITS = 4K: a text word's, tail form)
template <int BITS>
MRX_HD uint32_t push(uint32_t acc, uint32_t d) {
  if (BITS >= 32) return d;
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(d, acc, BITS);
#else
  return (acc << (BITS & 31)) | (d >> ((32 - BITS) & 31));
#endif
}
MRX_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
// highest set bit of an accumulator word -> (byte inside the word's group, slot); clears it
template <int K>
MRX_HD void pop_hit(uint32_t &m, int &byte, int &slot) {
  const int x = clz32(m);
  m &= ~(0x80000000u >> x);
  byte = x / K;
  slot = x % K;
}
template <int K> struct Geometry {
  static constexpr int kGroup = 32 / K;                   // text bytes per accumulator word
  static constexpr int kAccWords = (144 + kGroup - 1) / kGroup;  // per 144-byte chunk: 5 / 9 / 18
};

}  // namespace sand
}  // namespace mrx
