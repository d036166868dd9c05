// CPU harness for mojo_regex_b200/csrc/next/sep_core.hpp: emulates k_sep tile by tile with the kernel's own
// index arithmetic (T threads x 144 bytes per tile, three halfword bit streams with 2 + 3 halo halfwords, 224-bit
// windows read back from the streams) and lists the separators.  `threads` is a parameter so that small texts
// span several tiles.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../mojo_regex_b200/csrc/next/sep_core.hpp"
#include "../../mojo_regex_b200/csrc/next/sep_plan.hpp"

using namespace mrx::sep;

static void bits_at(const uint8_t *text, int64_t n, int64_t pos, const Classes &k, uint32_t &c, uint32_t &l, uint32_t &b) {
  // 16 bytes starting at pos (may start before the text or run past its end)
  uint8_t buf[16] = {0};
  c = l = b = 0;
  uint32_t valid = 0;
  for (int i = 0; i < 16; ++i)
    if (pos + i >= 0 && pos + i < n) { buf[i] = text[pos + i]; valid |= 1u << i; }
  uint32_t w[4];
  std::memcpy(w, buf, 16);
  bits16(w[0], w[1], w[2], w[3], 16, k, c, l, b);
  c &= valid; l &= valid; b &= valid;
}

static int64_t scan_with(const Classes &k, const uint8_t *text, int64_t n, int threads, int64_t *out, int64_t cap,
                         int *overflow);

extern "C" int64_t sep_scan(const uint8_t *text, int64_t n, int threads, int needle, int range_lo, int range_hi,
                            const int *extra, int n_extra, const int *bad, int n_bad, int64_t *out, int64_t cap,
                            int *overflow) {
  Classes k{};
  k.needle = (uint32_t)needle * 0x01010101u;
  k.range_lo = (uint32_t)range_lo;
  k.range_hi = (uint32_t)range_hi;
  k.n_extra = n_extra;
  for (int i = 0; i < n_extra; ++i) k.extra[i] = (uint32_t)extra[i] * 0x01010101u;
  k.n_bad = n_bad;
  for (int i = 0; i < n_bad; ++i) k.bad[i] = (uint32_t)bad[i] * 0x01010101u;
  return scan_with(k, text, n, threads, out, cap, overflow);
}

// planner + scan: -2 when the automaton is not of the X* c Y* shape (or its sets cannot be expressed)
extern "C" int64_t sep_detect_scan(const int32_t *trans, const uint8_t *accepting, int n_states, int start,
                                   const uint8_t *lookup, int c, const uint8_t *text, int64_t n, int threads,
                                   int64_t *out, int64_t cap, int *overflow) {
  Classes k{};
  if (!detect(trans, accepting, n_states, start, lookup, c, k)) return -2;
  return scan_with(k, text, n, threads, out, cap, overflow);
}

static int64_t scan_with(const Classes &k, const uint8_t *text, int64_t n, int threads, int64_t *out, int64_t cap,
                         int *overflow) {
  *overflow = 0;
  int64_t m = 0;
  const int64_t tile_bytes = (int64_t)threads * kChunk;
  const int hw = kStreamPre + 9 * threads + kStreamPost + 1;
  std::vector<uint16_t> bc((size_t)hw), bl((size_t)hw), bb((size_t)hw);
  for (int64_t tile_start = 0; tile_start < n; tile_start += tile_bytes) {
    uint32_t c, l, b;
    for (int tid = 0; tid < threads; ++tid)
      for (int v = 0; v < 9; ++v) {
        bits_at(text, n, tile_start + (int64_t)tid * kChunk + 16 * v, k, c, l, b);
        bc[(size_t)stream_chunk(tid, v)] = (uint16_t)c; bl[(size_t)stream_chunk(tid, v)] = (uint16_t)l; bb[(size_t)stream_chunk(tid, v)] = (uint16_t)b;
      }
    for (int i = 0; i < kStreamPre; ++i) {
      bits_at(text, n, tile_start - 32 + 16 * i, k, c, l, b);
      bc[(size_t)i] = (uint16_t)c; bl[(size_t)i] = (uint16_t)l; bb[(size_t)i] = (uint16_t)b;
    }
    for (int v = 0; v < kStreamPost; ++v) {
      bits_at(text, n, tile_start + tile_bytes + 16 * v, k, c, l, b);
      bc[(size_t)stream_post(threads, v)] = (uint16_t)c; bl[(size_t)stream_post(threads, v)] = (uint16_t)l; bb[(size_t)stream_post(threads, v)] = (uint16_t)b;
    }
    for (int tid = 0; tid < threads; ++tid) {
      Window w;
      window_from_streams(bc.data(), bl.data(), bb.data(), tid, w);
      const int64_t chunk0 = tile_start + (int64_t)tid * kChunk;
      for (int kk = 1; kk < kWords - 1; ++kk) {  // chunk = window positions 32 .. 175
        uint32_t f = w.c[kk];
        if (kk == 5) f &= 0xffffu;
        while (f) {
          const int j = __builtin_ctz(f);
          f &= f - 1;
          int back, fwd;
          if (separator_at(w.l[kk - 1], w.l[kk], w.l[kk + 1], w.bad[kk - 1], w.bad[kk], w.bad[kk + 1], j, &back, &fwd, overflow)) {
            const int64_t pos = chunk0 - kPre + 32 * kk + j;
            if (m < cap) { out[2 * m] = pos - back; out[2 * m + 1] = pos + 1 + fwd; }
            ++m;
          }
        }
      }
    }
  }
  return m;
}
