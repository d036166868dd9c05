// CPU harness for mojo_regex_b200/csrc/sand.hpp: scans a text the way k_sand's threads do -- tiles of 256
// chunks of 144 bytes, every chunk warmed up on `warm_words` 4-byte words of history (the tile's first chunk only
// on what is staged before the tile: min(32, tile_start) bytes), hit bits pushed into accumulator words of
// 32 / K bytes, the last partial word normalised, hits popped from the top (clz) and decoded as (byte, slot) --
// with the same helpers (step / push / pop_hit / Geometry) the kernel uses, and lists the occurrences.
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "../../mojo_regex_b200/csrc/sand.hpp"

using namespace mrx::sand;

template <int K>
static int64_t scan_k(const Table &t, const uint8_t *text, int64_t n, int64_t origin, int64_t key_lo, int64_t key_hi,
                      int64_t *out, int64_t cap) {
  constexpr int G = Geometry<K>::kGroup, kChunk = 144, kThreads = 256, kTile = kThreads * kChunk, kAcc = (kChunk + G - 1) / G;
  int64_t m = 0;
  for (int64_t tile_start = origin; tile_start < n; tile_start += kTile) {
    const int pre = (int)std::min<int64_t>(32, tile_start);
    for (int tid = 0; tid < kThreads; ++tid) {
      const int64_t c0 = tile_start + (int64_t)tid * kChunk;
      int ww = t.warm_words;
      if (tid == 0) ww = std::min(ww, pre >> 2);
      uint32_t D = 0;
      for (int i = ww; i > 0; --i)
        for (int k = 0; k < 4; ++k) {
          const int64_t q = c0 - 4 * i + k;
          D = step<K>(D, t.B[q < n ? text[q] : 0xAA], t.init);   // bytes past the text: stale buffer contents
        }
      uint32_t acc_words[kAcc];
      uint32_t acc = 0;
      for (int done = 1; done <= kChunk; ++done) {
        const int64_t q = c0 + done - 1;
        D = step<K>(D, t.B[q < n ? text[q] : 0xAA], t.init);
        if (!t.tail) acc = push<K>(acc, D);
        else if (done % 4 == 0) acc = push<4 * K>(acc, D);   // tail form: a word's hit bits are the top 4K bits of D
        if (done % 4 == 0 && (done % G == 0 || done == kChunk)) {
          if (done % G != 0) acc <<= K * (G - done % G);
          acc_words[(done - 1) / G] = acc;
          acc = 0;
        }
      }
      for (int a = 0; a < kAcc; ++a) {
        uint32_t mm = acc_words[a];
        while (mm) {
          int b, slot;
          pop_hit<K>(mm, b, slot);
          const int64_t end = c0 + a * G + b + 1;
          const int64_t start = end - (int64_t)((t.lens >> (8 * slot)) & 0xffu);
          if (end <= n && start >= key_lo && start < key_hi) {   // the kernel's edge-tile filter
            if (m < cap) { out[2 * m] = start; out[2 * m + 1] = end; }
            ++m;
          }
        }
      }
    }
  }
  return m;
}

extern "C" int64_t sand_scan(const uint8_t *text, int64_t n, const char *lits_nl, int64_t origin, int64_t key_lo,
                             int64_t key_hi, int64_t *out, int64_t cap, int allow_tail) {
  std::vector<std::string> lits;
  std::string cur;
  for (const char *p = lits_nl; *p; ++p) {
    if (*p == '\n') { lits.push_back(cur); cur.clear(); } else cur.push_back(*p);
  }
  if (!cur.empty()) lits.push_back(cur);
  Table t;
  if (!build(lits, t, allow_tail != 0)) return -1;
  if (allow_tail == 2 && !t.tail) return -2;   // the caller wanted to see the tail form
  if (t.k == 1) return scan_k<1>(t, text, n, origin, key_lo, key_hi, out, cap);
  if (t.k == 2) return scan_k<2>(t, text, n, origin, key_lo, key_hi, out, cap);
  return scan_k<4>(t, text, n, origin, key_lo, key_hi, out, cap);
}
