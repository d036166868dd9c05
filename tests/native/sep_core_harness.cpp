// CPU harness for mojo_regex_b200/csrc/next/sep_core.hpp: walks a text the way the kernel's threads will
// (144-byte chunks, 224-bit windows built from SWAR membership flags) and lists the separators.
#include <cstdint>
#include <cstring>

#include "../../mojo_regex_b200/csrc/next/sep_core.hpp"

using namespace mrx::sep;

extern "C" int64_t sep_scan(const uint8_t *text, int64_t n, int needle, int range_lo, int range_hi, const int *extra,
                            int n_extra, const int *bad, int n_bad, int64_t *out, int64_t cap, int *overflow) {
  Classes k{};
  k.needle = (uint32_t)needle * 0x01010101u;
  k.range_lo = (uint32_t)range_lo;
  k.range_hi = (uint32_t)range_hi;
  k.n_extra = n_extra;
  for (int i = 0; i < n_extra; ++i) k.extra[i] = (uint32_t)extra[i] * 0x01010101u;
  k.n_bad = n_bad;
  for (int i = 0; i < n_bad; ++i) k.bad[i] = (uint32_t)bad[i] * 0x01010101u;
  *overflow = 0;
  int64_t m = 0;
  for (int64_t chunk0 = 0; chunk0 < n; chunk0 += kChunk) {
    Window w{};
    for (int word = 0; word < kWords; ++word) {
      for (int q = 0; q < 8; ++q) {  // 8 groups of 4 bytes per 32-bit word of the window
        const int64_t pos = chunk0 - kPre + 32 * word + 4 * q;
        uint8_t four[4] = {0, 0, 0, 0};
        uint32_t valid = 0;
        for (int b = 0; b < 4; ++b)
          if (pos + b >= 0 && pos + b < n) { four[b] = text[pos + b]; valid |= 1u << b; }
        uint32_t packed;
        std::memcpy(&packed, four, 4);
        const WordBits r = classify4(packed, k);
        w.c[word] |= (r.c & valid) << (4 * q);
        w.l[word] |= (r.l & valid) << (4 * q);
        w.bad[word] |= (r.bad & valid) << (4 * q);
      }
    }
    for (int p = kPre; p < kPre + kChunk; ++p) {
      if (!((w.c[p >> 5] >> (p & 31)) & 1u)) continue;
      int s, e;
      if (separator_at(w, p, &s, &e, overflow)) {
        if (m < cap) { out[2 * m] = chunk0 - kPre + s; out[2 * m + 1] = chunk0 - kPre + e; }
        ++m;
      }
    }
  }
  return m;
}
