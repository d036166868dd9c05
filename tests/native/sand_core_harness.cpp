// CPU harness for mojo_regex_b200/csrc/next/sand_core.hpp: scans a text the way the kernel's threads will
// (144-byte chunks, each warmed up on the max_len - 1 bytes before it, top bytes packed four to a word, hits
// gathered from the packed words) and lists the occurrences.
#include <cstdint>
#include <string>
#include <vector>

#include "../../mojo_regex_b200/csrc/next/sand_core.hpp"

using namespace mrx::sand;

extern "C" int64_t sand_scan(const uint8_t *text, int64_t n, const char *lits_nl, int64_t *out, int64_t cap) {
  std::vector<std::string> lits;
  std::string cur;
  for (const char *p = lits_nl; *p; ++p) {
    if (*p == '\n') { lits.push_back(cur); cur.clear(); } else cur.push_back(*p);
  }
  if (!cur.empty()) lits.push_back(cur);
  Table t;
  if (!build(lits, t)) return -1;
  int64_t m = 0;
  constexpr int kChunk = 144;
  for (int64_t chunk0 = 0; chunk0 < n; chunk0 += kChunk) {
    uint32_t d = 0;
    for (int64_t q = chunk0 - (t.max_len - 1) < 0 ? 0 : chunk0 - (t.max_len - 1); q < chunk0; ++q)
      d = step(d, text[q], t.B, t.init, t.s);
    for (int64_t w0 = chunk0; w0 < chunk0 + kChunk && w0 < n; w0 += 4) {
      uint32_t packed = 0;
      int live = 0;
      for (int j = 0; j < 4 && w0 + j < n; ++j, ++live) {
        d = step(d, text[w0 + j], t.B, t.init, t.s);
        packed |= (d >> 24) << (8 * j);
      }
      uint32_t hits = hit_nibble(packed, t.acc4) & ((1u << live) - 1u);
      while (hits) {
        const int j = __builtin_ctz(hits);
        hits &= hits - 1;
        const uint32_t top = (packed >> (8 * j)) & 0xffu & (t.acc4 & 0xffu);
        const int i = literal_of(top);
        const int64_t end = w0 + j + 1;
        if (m < cap) { out[2 * m] = end - t.len[i]; out[2 * m + 1] = end; }
        ++m;
      }
    }
  }
  return m;
}
