// k_sep -- finder for patterns of the shape X* c Y* (C5: `\s*,\s*`) on membership bits.  EXPERIMENTAL: written
// at the end of round 1 with no GPU time left, compiled for sm_100a by tests/test_next_kernels_compile_cpu.py
// but never launched; build.py does not put it into libmojoregex_cuda.so.  The arithmetic it relies on
// (sep_core.hpp) is verified against the oracle on the CPU; the plumbing below (staging, ranks, look-back, span
// emission) follows k_runs' CLASS_RUNS path and still has to see a GPU.  DESIGN.md section 8, item 1.
//
// Reference semantics: HybridMatcher._match_all_required_byte (matcher.mojo:864-898) for an automaton
// S0 -X-> S1 -c-> S2(acc) -Y-> S3(acc) (dfa.mojo:699-807): one span per occurrence of c,
// start = c - (run of back-scan bytes before it), none when that run holds a byte X lacks, end = c + 1 + (run
// of Y bytes after it).  Spans may overlap (two commas sharing one blank run); split then reports
// MRX_E_UNDEFINED as it does today.
//
// One CTA per 36 864-byte tile, thread t owns bytes [144 t, 144 t + 144).
//   1. tile + 256-byte halo staged by one TMA bulk copy;
//   2. every thread turns its 9 vectors into 9 halfwords of membership bits per set (comma / back-scan set /
//      bad bytes) with SWAR flags -- no table lookups -- and parks them in three bit streams in shared memory;
//      five more threads do the 32 bytes before the tile (from HBM) and the 48 after it (from the halo);
//   3. every thread reads the 14 halfwords (224 bits) around its chunk and visits its commas: count-leading-
//      ones / count-trailing-ones on 32-bit windows give the run lengths;
//   4. counts -> block scan -> decoupled look-back -> spans written in text order.
// A blank run of 32 or more bytes next to a comma cannot be measured inside the window: the kernel raises
// *overflow and the host repeats the call with k_scan.
#include "../device.cuh"
#include "sep_core.hpp"

namespace mrx {

struct SepParams {
  const uint8_t *text;
  int64_t n;
  int64_t tile_origin;
  int64_t key_lo, key_hi;  // commas reported: [key_lo, key_hi)
  sep::Classes cls;
  int32_t n_tiles;
  int32_t tile_base;
  int64_t *spans;
  int64_t cap;
  int64_t rebase;
  unsigned long long *desc;
  unsigned int *ticket;
  unsigned long long *total;
  int *overflow;
};

constexpr int kSepHw = kThreads * 9;  // halfwords of one tile's bit stream
struct SepShared {
  alignas(16) uint8_t text[kTile + kHalo];
  alignas(8) uint64_t bar;
  int scan_ws[kThreads / 32 + 2];
  int tile;
  unsigned long long base;
  // bit streams: [0, 2) = the 32 bytes before the tile, [2, 2 + kSepHw) = the tile, then 3 halfwords of halo (+1 pad)
  uint16_t bc[kSepHw + 6], bl[kSepHw + 6], bb[kSepHw + 6];
};

__device__ __forceinline__ void sep_bits16(const uint4 &v, int live, const sep::Classes &k, uint32_t &c, uint32_t &l,
                                           uint32_t &b) {
  sep::bits16(v.x, v.y, v.z, v.w, live, k, c, l, b);
}

__global__ void __launch_bounds__(kThreads, 4) k_sep(const SepParams prm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SepShared &sh = *reinterpret_cast<SepShared *>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    sh.tile = prm.tile_base + (int)atomicAdd(prm.ticket, 1u);
    mbar_init(&sh.bar, 1);
    fence_proxy_async();
  }
  __syncthreads();
  const int tile = sh.tile;
  const int64_t tile_start = prm.tile_origin + (int64_t)tile * kTile;
  int stage_bytes = kTile + kHalo;
  if (tile_start + stage_bytes > prm.n) stage_bytes = (int)(prm.n - tile_start > 0 ? prm.n - tile_start : 0);
  {  // stage the tile and its halo
    const bool aligned = (reinterpret_cast<uintptr_t>(prm.text + tile_start) & 15) == 0;
    const int bulk = aligned ? (stage_bytes & ~15) : 0;
    if (tid == 0 && bulk > 0) {
      mbar_expect_tx(&sh.bar, (uint32_t)bulk);
      tma_bulk_g2s(sh.text, prm.text + tile_start, (uint32_t)bulk, &sh.bar);
    }
    for (int i = bulk + tid; i < stage_bytes; i += kThreads) sh.text[i] = prm.text[tile_start + i];
    if (bulk > 0) mbar_wait(&sh.bar, 0);
    __syncthreads();
  }
  // ---- membership bits of my chunk -> the three bit streams
  const int chunk_rel = tid * kChunk;
#pragma unroll
  for (int v = 0; v < kVecs; ++v) {
    const uint4 w = *reinterpret_cast<const uint4 *>(sh.text + chunk_rel + 16 * v);
    uint32_t c, l, b;
    sep_bits16(w, stage_bytes - (chunk_rel + 16 * v), prm.cls, c, l, b);
    sh.bc[sep::stream_chunk(tid, v)] = (uint16_t)c;
    sh.bl[sep::stream_chunk(tid, v)] = (uint16_t)l;
    sh.bb[sep::stream_chunk(tid, v)] = (uint16_t)b;
  }
  if (tid < 2) {  // the 32 bytes before the tile, 16 per thread, straight from HBM (bytes before the text: no bits)
    const int64_t p0 = tile_start - 32 + 16 * tid;
    uint32_t c = 0, l = 0, b = 0;
    for (int q = 0; q < 4; ++q) {
      uint32_t packed = 0, valid = 0;
      for (int j = 0; j < 4; ++j) {
        const int64_t p = p0 + 4 * q + j;
        if (p >= 0 && p < prm.n) { packed |= (uint32_t)prm.text[p] << (8 * j); valid |= 1u << j; }
      }
      const sep::WordBits r = sep::classify4(packed, prm.cls);
      c |= (r.c & valid) << (4 * q);
      l |= (r.l & valid) << (4 * q);
      b |= (r.bad & valid) << (4 * q);
    }
    sh.bc[tid] = (uint16_t)c; sh.bl[tid] = (uint16_t)l; sh.bb[tid] = (uint16_t)b;
  } else if (tid < 5) {  // the 48 bytes after the tile, from the staged halo
    const int v = tid - 2;
    const uint4 w = *reinterpret_cast<const uint4 *>(sh.text + kTile + 16 * v);
    uint32_t c, l, b;
    sep_bits16(w, stage_bytes - (kTile + 16 * v), prm.cls, c, l, b);
    const int h = sep::stream_post(kThreads, v);
    sh.bc[h] = (uint16_t)c; sh.bl[h] = (uint16_t)l; sh.bb[h] = (uint16_t)b;
  }
  __syncthreads();
  // ---- my window: 14 halfwords = 224 bits starting 32 bits before my chunk
  sep::Window w;
  sep::window_from_streams(sh.bc, sh.bl, sh.bb, tid, w);
  // ---- pass 1: which of my commas carry a separator (bit j of ok[k] <-> window position 32 k + j)
  const int64_t chunk0 = tile_start + chunk_rel;
  const bool edge_tile = tile_start < prm.key_lo || tile_start + kTile > prm.key_hi;
  uint32_t ok[sep::kWords - 1];
  int cnt = 0, ovf = 0;
#pragma unroll
  for (int k = 1; k < sep::kWords - 1; ++k) {  // chunk = window positions 32 .. 175: words 1..4 and half of word 5
    uint32_t f = w.c[k];
    if (k == 5) f &= 0xffffu;
    uint32_t good = 0;
    while (f) {
      const int j = __ffs((int)f) - 1;
      f &= f - 1;
      if (edge_tile) {
        const int64_t key = chunk0 - sep::kPre + 32 * k + j;
        if (key < prm.key_lo || key >= prm.key_hi) continue;
      }
      int back, fwd;
      if (sep::separator_at(w.l[k - 1], w.l[k], w.l[k + 1], w.bad[k - 1], w.bad[k], w.bad[k + 1], j, &back, &fwd, &ovf)) {
        good |= 1u << j;
        ++cnt;
      }
    }
    ok[k] = good;
  }
  if (ovf) atomicOr(prm.overflow, 1);
  // ---- ranks
  int total;
  const int rank = block_exclusive_scan(cnt, sh.scan_ws, &total);
  if (tid < 32) {
    const unsigned long long b = lookback_exclusive(prm.desc, tile, (unsigned long long)total);
    if (tid == 0) {
      sh.base = b;
      if (tile == prm.n_tiles - 1) *prm.total = b + (unsigned long long)total;
    }
  }
  __syncthreads();
  // ---- pass 2: emit
  unsigned long long o = sh.base + (unsigned long long)rank;
#pragma unroll
  for (int k = 1; k < sep::kWords - 1; ++k) {
    uint32_t f = ok[k];
    while (f) {
      const int j = __ffs((int)f) - 1;
      f &= f - 1;
      int back, fwd, dummy = 0;
      sep::separator_at(w.l[k - 1], w.l[k], w.l[k + 1], w.bad[k - 1], w.bad[k], w.bad[k + 1], j, &back, &fwd, &dummy);
      const int64_t pos = chunk0 - sep::kPre + 32 * k + j;  // the comma
      if (o < (unsigned long long)prm.cap) st_span(prm.spans + 2 * o, pos - back + prm.rebase, pos + 1 + fwd + prm.rebase);
      ++o;
    }
  }
}

// grid = tiles handed out by this launch; smem = sizeof(SepShared)
void launch_sep(const SepParams &p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_sep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SepShared));
    attr_set = true;
  }
  k_sep<<<grid, kThreads, sizeof(SepShared), st>>>(p);
}

}  // namespace mrx
