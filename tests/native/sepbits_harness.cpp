// CPU harness for mojo_regex_b200/csrc/sepbits.hpp: splits a text chunk by chunk exactly as the threads of k_split
// do -- 144-byte chunks, a 288-bit window per chunk (96 bits before it, 48 after it; bits of bytes outside the text
// are 0), sepb::separators for the kept / separator masks of the chunk -- then compacts the kept bytes and records,
// for every valid separator byte, the number of kept bytes before it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../mojo_regex_b200/csrc/sepbits.hpp"

using namespace mrx::sepb;

// returns the number of pieces; flags: bit 0 = overflow (two adjacent 16-byte groups, aligned to 16, entirely in L),
// bit 1 = overlapping separators
extern "C" int64_t sepb_split(const uint8_t *text, int64_t n, const uint32_t *table, uint8_t *out_bytes, int64_t *out_offs,
                              int *flags) {
  constexpr int kChunk = 144;
  *flags = 0;
  int64_t n_out = 0, n_sep = 0;
  out_offs[0] = 0;
  auto cls = [&](int64_t q, int shift) -> uint32_t { return (q >= 0 && q < n) ? (table[text[q]] >> shift) & 1u : 0u; };
  // overflow rule of the kernel
  for (int64_t g = 0; g + 32 <= ((n + 15) / 16) * 16; g += 16) {
    bool full = true;
    for (int64_t q = g; q < g + 32 && full; ++q) full = cls(q, 8) != 0;
    if (full) *flags |= 1;
  }
  for (int64_t c0 = 0; c0 < n; c0 += kChunk) {
    Win C, L, B;
    for (int k = 0; k < kWin; ++k) {
      C.w[k] = L.w[k] = B.w[k] = 0;
      for (int j = 0; j < 32; ++j) {
        const int64_t q = c0 - kPreBits + 32 * k + j;
        C.w[k] |= cls(q, 0) << j;
        L.w[k] |= cls(q, 8) << j;
        B.w[k] |= cls(q, 16) << j;
      }
    }
    const Own o = separators(C, L, B);
    if (o.overlap) *flags |= 2;
    for (int i = 0; i < kChunk && c0 + i < n; ++i) {
      if ((o.sep[i >> 5] >> (i & 31)) & 1u) out_offs[++n_sep] = n_out;
      if ((o.kept[i >> 5] >> (i & 31)) & 1u) out_bytes[n_out++] = text[c0 + i];
    }
  }
  out_offs[n_sep + 1] = n_out;
  return n_sep + 1;
}
