// Core of the Shift-And formulation of the literal-set scan (round-2 candidate for k_dense, DESIGN.md section 8,
// item 2a) -- NOT part of the library yet: no product source includes this header.  Checked on the CPU by
// tests/test_sand_core_cpu.py (g++ harness against the oracle).
//
// k literals, stride s = k.  Literal i keeps its LAST state in bit 31 - i and the state j characters before its
// end in bit 31 - i - s*j.  One step per text byte:   D = ((D << s) + INIT) & B[byte]
//   * a shift by s moves every literal's states one character forward; bits shifted past bit 31 are the states
//     after the last one and drop out;
//   * INIT holds the first-state bit of every literal; no shifted bit can land on one (same residue mod s =
//     same literal, and there is no state before the first), so OR and ADD coincide: one IMAD;
//   * "literal i ends at this byte" <=> bit 31 - i of D, i.e. the top k bits of D's TOP BYTE (k <= 8) -- the
//     byte the kernel packs per text byte, exactly where k_dense keeps its state byte today.
// The lookup address depends on the text only (no dependent chain) and B has 4-byte entries: the bytes of
// ordinary text fall in different banks.
// Preconditions (the caller's: program.cpp build_dense already enforces them for the dense path): no literal
// is a proper prefix of another or occurs inside another, so at most one literal ends at any byte.
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

struct Table {
  int32_t k = 0, s = 0, max_len = 0;
  uint32_t init = 0;
  uint32_t acc4 = 0;     // accept bits of the top byte (its top k bits), in all four bytes of a word
  uint8_t len[8] = {0};  // literal length by literal index
  uint32_t B[256] = {0};
};

inline bool build(const std::vector<std::string> &lits, Table &t) {
  t = Table();
  const int k = (int)lits.size();
  if (k < 1 || k > 8) return false;
  size_t max_len = 0;
  for (const std::string &w : lits) {
    if (w.empty() || w.size() > 255) return false;
    max_len = w.size() > max_len ? w.size() : max_len;
  }
  if ((size_t)k * (max_len - 1) + (size_t)k > 32) return false;
  t.k = k; t.s = k; t.max_len = (int)max_len;
  for (int i = 0; i < k; ++i) {
    const std::string &w = lits[(size_t)i];
    const int n = (int)w.size();
    for (int j = 0; j < n; ++j) t.B[(unsigned char)w[(size_t)(n - 1 - j)]] |= 1u << (31 - i - t.s * j);
    t.init |= 1u << (31 - i - t.s * (n - 1));
    t.len[i] = (uint8_t)n;
  }
  t.acc4 = ((0xffu << (8 - k)) & 0xffu) * 0x01010101u;
  return true;
}

MRX_HD uint32_t step(uint32_t d, uint32_t byte, const uint32_t *B, uint32_t init, int s) {
  return ((d << s) + init) & B[byte];
}

// four packed top bytes -> 4 bits: byte j holds an accept bit (byte 0 -> bit 0); exact, no cross-byte carries
MRX_HD uint32_t hit_nibble(uint32_t packed, uint32_t acc4) {
  const uint32_t t = packed & acc4;
  const uint32_t nz = (((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
  return (((nz >> 7) * 0x01020408u) >> 24) & 0xfu;
}

// index of the literal whose accept bit is set in a packed top byte (bit 7 <-> literal 0)
MRX_HD int literal_of(uint32_t top_byte) {
#if defined(__CUDA_ARCH__)
  return __clz((int)(top_byte << 24));
#else
  return __builtin_clz(top_byte << 24);
#endif
}

}  // namespace sand
}  // namespace mrx
