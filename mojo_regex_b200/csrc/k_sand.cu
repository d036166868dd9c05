// k_sand<K, TAIL> -- findall for small literal sets (BASELINE config 2) as a bit-parallel Shift-And scan; see
// sand.hpp for the automaton.  Replaces k_dense wherever the set fits 32 state bits (K * max_len <= 32, K <= 4
// slots); k_dense stays for larger sets and for search.
//
// What bounded k_dense (profiles/r01_k_dense_h_ncu.txt: LSU wavefronts 79 % of peak, 2.04 wavefronts per table
// lookup, 121 M wavefronts per GiB of which only 68 M were the walk) and what this kernel does about it:
//   * the lookup address depends on the text byte only (B[byte], 4-byte entries): no dependent chain of
//     shared-memory loads, and ordinary text spreads over the banks -- one wavefront per lookup;
//   * hits are not extracted byte by byte: the hit bits of the state are funnel-shifted into an accumulator (one
//     SHF per byte; one per 4-byte word in the tail form); a 144-byte chunk leaves 5 / 9 / 18 accumulator words,
//     parked in shared memory [word][thread] (conflict free), counted with POPC and decoded one tile later,
//     straight from the bit positions (byte = clz / K, literal = clz % K) -- no state bytes written back over the
//     text, no per-hit table loads, no private token lists, no recompute path;
//   * the warm-up runs over exactly ceil((max_len - 1) / 4) words of history instead of 16 predicated steps;
//   * the text is double buffered: the TMA bulk copy of tile i + 1 is issued BEFORE tile i is walked;
//   * the integer work is split between the ALU pipe (LOP3 / SHF) and the FMA pipe (IMAD, IMAD.HI): each pipe
//     starts one warp instruction every 2 cycles per scheduler, so a loop that keeps everything on the ALU pipe
//     runs at half the issue rate.  Per byte: IMAD.HI (shift of the address), LOP3 (its mask), LDS, IMAD
//     ((D << K) + INIT), LOP3 (the AND) [+ SHF, the push, when the set does not fit the tail form].
// Persistent CTAs; tiles are handed out by an atomic ticket (SMs do not all run at the same speed: with a static
// assignment the fast CTAs spent 8 % of the kernel's instructions polling for the slow ones' counts,
// profiles/r02_k_sand_c_ncu.txt), fetched one iteration ahead so that its latency hides behind a walk.  Ordered
// output: every CTA keeps the exclusive prefix of its previous tile and adds the aggregates of the tiles in
// between (a sliding window over the descriptor array, about gridDim entries); a tile's matches wait as 16-bit
// tokens in per-warp rings and are written out kSandLag tiles later, the window's descriptor loads are issued
// before a walk and consumed after it: nobody waits for a predecessor that is still scanning.  Deadlock free
// because every CTA is resident and a tile's aggregate is published before its owner waits for anything.
#include <cstdlib>

#include "kernels.cuh"
#include "sand.hpp"

namespace mrx {

// Geometry: 256 threads x 144 bytes (9 x LDS.128; 9 is odd, so a quarter-warp's vectors hit 8 distinct bank groups),
// two CTAs per SM.  Smaller chunks were measured (512 threads x 80 bytes, 32 warps per SM: 0.62 ms per GiB against
// 0.50): the per-thread work of a tile that does not depend on the chunk length (scan, window, sums, warm-up) costs
// more than the extra warps hide.  Three chains of 48 bytes per thread instead of two (64 + 80) were measured too:
// 0.481 ms against 0.471 -- the third warm-up costs more than the extra chain in flight gains, the walk is not
// latency-bound.
constexpr int kSandThreads = 256;
constexpr int kSandChunk = kSandTileBytes / kSandThreads;  // 144
constexpr int kSandVecs = kSandChunk / 16;                 // 9
constexpr int kSandSplit = 64;   // a thread walks its chunk as two independent chains: bytes [0, 64) and [64, 144)
constexpr int kSandPre = 32;     // bytes of left context staged before a tile (history of the tile's first chunk)
constexpr int kSandLag = 2;      // a tile's matches are written out this many tiles after it was walked
constexpr int kSandRing = 512;   // tokens per warp waiting to be written out (power of two)
static_assert(kSandChunk * kSandThreads == kSandTileBytes && kSandChunk % 16 == 0 && (kSandVecs & 1) == 1, "geometry");

template <int K>
struct SandShared {
  static constexpr int kAcc = (kSandChunk + sand::Geometry<K>::kGroup - 1) / sand::Geometry<K>::kGroup;
  alignas(128) uint8_t text[2][kSandPre + kSandTileBytes];
  uint32_t acc[kAcc][kSandThreads];  // hit bits of the chunks of the tile just walked, [word][thread]
  uint32_t B[256];
  alignas(8) uint64_t bar[2];
  unsigned long long red_ws[3][kSandThreads / 32];  // window sums per warp: two generations + the blocking path's
  int scan_ws[2][kSandThreads / 32];                // hit counts per warp, two generations (one barrier per tile)
  uint16_t ring[kSandThreads / 32][kSandRing];      // per warp FIFO: (position in the warp's 4608 bytes) << 2 | slot
  int next_tile[2];                                 // ticket broadcast: the tile of the iteration after next
};

// B[byte k of w].  Three ways to form the byte offset 4 * byte, which differ in the pipes they load (each of the
// ALU and FMA pipes starts one warp instruction every two cycles per scheduler):
//   ADDR 0:  SHF + LOP3            (w >> (8k - 2)) & 0x3fc                      both on the ALU pipe
//   ADDR 1:  IMAD.HI + LOP3        umulhi(w, 2^(34 - 8k)) & 0x3fc               shift on the FMA pipe
//   ADDR 2:  PRMT + IMAD           byte_perm(w, 0, 0x4440 + k) * 4              extract on ALU, scale on FMA
template <int ADDR>
__device__ __forceinline__ uint32_t sand_entry(const uint32_t *B, uint32_t w, int k) {
  uint32_t off;
  if (ADDR == 0) {
    off = (k == 0 ? w << 2 : w >> (8 * k - 2)) & 0x3fcu;
  } else if (ADDR == 1) {
    off = (k == 0 ? w * 4u : __umulhi(w, 1u << (32 - (8 * k - 2)))) & 0x3fcu;
  } else {
    const uint32_t b = __byte_perm(w, 0, 0x4440 + k);
    asm("mad.lo.u32 %0, %1, 4, 0;" : "=r"(off) : "r"(b));
  }
  return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(B) + off);
}

template <int K, int ADDR>
__device__ __forceinline__ uint32_t sand_warm(const uint32_t *B, uint32_t init, const uint8_t *at, int words) {
  uint32_t D = 0;
  for (int i = words; i > 0; --i) {
    const uint32_t w = *reinterpret_cast<const uint32_t *>(at - 4 * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) D = sand::step<K>(D, sand_entry<ADDR>(B, w, k), init);
  }
  return D;
}

// One thread's 144 bytes as two independent chains (A: bytes [0, 64), B: [64, 144), both warmed up by the caller),
// interleaved instruction by instruction so that a warp always has two dependent IMAD -> LOP3 chains in flight.
// Accumulator words go to shared memory (stride kSandThreads); returns the hit count, nz = which words are non-zero.
template <int K, bool TAIL, int ADDR>
__device__ __forceinline__ int sand_walk(const uint32_t *B, uint32_t init, const uint8_t *chunk, uint32_t Da, uint32_t Db,
                                         uint32_t *acc_out, uint32_t &nz_out) {
  constexpr int G = sand::Geometry<K>::kGroup;
  static_assert(kSandSplit % G == 0, "chain B must start on an accumulator word");
  int cnt = 0;
  uint32_t acc_a = 0, acc_b = 0, nz = 0;
  auto word = [&](uint32_t w, uint32_t &D, uint32_t &acc, int done /* bytes of the chunk consumed after this word */) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      D = sand::step<K>(D, sand_entry<ADDR>(B, w, k), init);
      if (!TAIL) acc = sand::push<K>(acc, D);
    }
    if (TAIL) acc = sand::push<4 * K>(acc, D);  // the word's hit bits sit in the top 4K bits of D
    if (done % G == 0 || done == kSandChunk) {
      if (done % G != 0) acc <<= K * (G - done % G);  // the last, partial group (K = 1): first byte on top as well
      const int a = (done - 1) / G;
      acc_out[a * kSandThreads] = acc;
      cnt += __popc(acc);
      nz |= min(acc, 1u) << a;
      acc = 0;
    }
  };
#pragma unroll
  for (int v = 0; v < 5; ++v) {  // chain A has 4 vectors, chain B 5
    uint4 ta = make_uint4(0, 0, 0, 0);
    if (v < 4) ta = *reinterpret_cast<const uint4 *>(chunk + 16 * v);
    const uint4 tb = *reinterpret_cast<const uint4 *>(chunk + kSandSplit + 16 * v);
    const uint32_t wa[4] = {ta.x, ta.y, ta.z, ta.w}, wb[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
      if (v < 4) word(wa[wi], Da, acc_a, 16 * v + 4 * wi + 4);
      word(wb[wi], Db, acc_b, kSandSplit + 16 * v + 4 * wi + 4);
    }
  }
  nz_out = nz;
  return cnt;
}

template <int K, bool TAIL, int ADDR>
__global__ void __launch_bounds__(kSandThreads, 2) k_sand(const SandParams prm) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SandShared<K> &sh = *reinterpret_cast<SandShared<K> *>(smem_raw);
  constexpr int G = sand::Geometry<K>::kGroup;
  constexpr int kAcc = SandShared<K>::kAcc;
  constexpr int W = kSandThreads / 32;
  constexpr int L = kSandLag, R = kSandRing;
  constexpr int kTileB = kSandTileBytes, kChunkB = kSandChunk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool text_aligned = (reinterpret_cast<uintptr_t>(prm.text + prm.tile_origin) & 15) == 0;
  if (tid == 0) {
    mbar_init(&sh.bar[0], 1);
    mbar_init(&sh.bar[1], 1);
    fence_proxy_async();
  }
  const sand::Table *tb = prm.table;
  for (int i = tid; i < 256; i += kSandThreads) sh.B[i] = tb->B[i];
  const uint32_t init = tb->init, lens = tb->lens;
  const int warm_words = tb->warm_words;
  __syncthreads();

  // stage tile t into buffer s: [tile_start - pre, tile_start + tile) clipped to the text, pre = 0 / 16 / 32
  auto issue_load = [&](int t, int s) {  // thread 0 only
    if (t >= prm.n_tiles) return;
    const int64_t ts = prm.tile_origin + (int64_t)t * kTileB;
    const int pre = (int)min((int64_t)kSandPre, ts);
    int64_t bytes = min((int64_t)kTileB, prm.n - ts) + pre;
    if (bytes < 0) bytes = 0;
    const uint32_t bulk = text_aligned ? ((uint32_t)bytes & ~15u) : 0u;
    mbar_expect_tx(&sh.bar[s], bulk);
    if (bulk) tma_bulk_g2s(sh.text[s] + (kSandPre - pre), prm.text + ts - pre, bulk, &sh.bar[s]);
  };

  // ---- ordered output.  cta_base = matches before tile `covered` (everything this CTA has written out so far and
  //      every tile before that); the oldest pending tile p adds the aggregates of [covered, p), covered = this CTA's
  //      previous tile + 1: about gridDim descriptors.  The 2 * kSandThreads nearest to p are loaded early (two per
  //      thread); a longer window (this CTA fell far behind) is finished by window_rest.
  unsigned long long cta_base = prm.base0;
  int covered = prm.tile_base;
  int depth = 0;                                     // pending tiles (walked, tokens in the rings, not written out)
  int q_tile[L], q_total[L], q_wfirst[L], q_wtotal[L];  // oldest first; q_w*: this warp's first rank / match count
#pragma unroll
  for (int i = 0; i < L; ++i) q_tile[i] = q_total[i] = q_wfirst[i] = q_wtotal[i] = 0;
  uint32_t ring_head = 0, ring_tail = 0;             // this warp's FIFO (free-running)
  uint16_t *ring = sh.ring[warp];
  unsigned long long win[2] = {kFlagAgg, kFlagAgg};
  auto window_load = [&](int p) {  // issue the loads of this thread's descriptors of the window of tile p: [covered, p)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = p - 1 - (tid + kSandThreads * u);
      win[u] = i >= covered ? ld_volatile_u64(prm.desc + i) : kFlagAgg;  // outside the window: published, value 0
    }
  };
  auto window_sum = [&](int p) -> unsigned long long {  // this warp's part of the window (waits for stragglers)
    unsigned long long sum = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = p - 1 - (tid + kSandThreads * u);
      while ((win[u] >> 62) == 0) win[u] = ld_volatile_u64(prm.desc + i);  // (short waits; a back-off measured no gain)
      sum += win[u] & kValMask;
    }
    for (int i = p - 1 - 2 * kSandThreads - tid; i >= covered; i -= kSandThreads) {  // (rare) the rest of a long window
      unsigned long long d;
      do { d = ld_volatile_u64(prm.desc + i); } while ((d >> 62) == 0);
      sum += d & kValMask;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    return sum;
  };
  // write out `n` matches of this warp from the front of its ring: one lane per match, consecutive lanes write
  // consecutive 16-byte spans.  o = output index of the first, wbase = text position of the warp's first byte.
  auto emit = [&](unsigned long long o, int64_t wbase, uint32_t from, int n) {
    for (int j = lane; j < n; j += 32) {
      const uint32_t tok = ring[(from + (uint32_t)j) & (R - 1)];
      const int64_t end = wbase + (tok >> 2) + 1;
      const int64_t len = (int64_t)((lens >> (8 * (tok & 3u))) & 0xffu);
      const unsigned long long oo = o + (unsigned long long)j;
      if (oo < (unsigned long long)prm.cap) st_span(prm.spans + 2 * oo, end - len, end);
    }
  };
  auto warp_text_base = [&](int t) -> int64_t {
    return prm.tile_origin + (int64_t)t * kTileB + (int64_t)(warp * 32) * kChunkB + prm.rebase;
  };
  auto flush_oldest = [&](unsigned long long base) {  // base = matches before the oldest pending tile
    cta_base = base + (unsigned long long)q_total[0];
    covered = q_tile[0] + 1;
    if (tid == 0 && q_tile[0] == prm.n_tiles - 1) *prm.total = cta_base;
    __syncwarp();  // the tokens were written by other lanes of this warp
    emit(base + (unsigned long long)q_wfirst[0], warp_text_base(q_tile[0]), ring_head, q_wtotal[0]);
    ring_head += (uint32_t)q_wtotal[0];
#pragma unroll
    for (int i = 0; i + 1 < L; ++i) { q_tile[i] = q_tile[i + 1]; q_total[i] = q_total[i + 1]; q_wfirst[i] = q_wfirst[i + 1]; q_wtotal[i] = q_wtotal[i + 1]; }
    --depth;
    __syncwarp();  // ring entries are free again
  };
  auto block_base = [&](unsigned long long wsum) -> unsigned long long {  // cta_base + the window (two barriers)
    __syncthreads();
    if (lane == 0) sh.red_ws[2][warp] = wsum;
    __syncthreads();
    unsigned long long base = cta_base;
#pragma unroll
    for (int w = 0; w < W; ++w) base += sh.red_ws[2][w];
    return base;
  };
  auto flush_oldest_blocking = [&]() {  // off the fast path: nothing hides the descriptor loads
    window_load(q_tile[0]);
    flush_oldest(block_base(window_sum(q_tile[0])));
  };

  // tickets: the first two tiles now, then one per iteration, two iterations ahead of its walk (thread 0 holds the
  // atomic's result in a register across the walk and broadcasts it before the barrier)
  const bool fixed = (prm.debug & 4) != 0;  // measurement: static round-robin assignment instead of tickets
  int fixed_next = prm.tile_base + (int)blockIdx.x;
  auto take = [&]() -> int {  // thread 0 only
    if (fixed) { const int t = fixed_next; fixed_next += (int)gridDim.x; return t; }
    return prm.tile_base + (int)atomicAdd(prm.ticket, 1u);
  };
  if (tid == 0) {
    sh.next_tile[0] = take();
    sh.next_tile[1] = take();
    issue_load(sh.next_tile[0], 0);
  }
  __syncthreads();
  int tile = sh.next_tile[0], tile_next = sh.next_tile[1];
  for (int it = 0; tile < prm.n_tiles; ++it) {
    const int s = it & 1;
    // buffer s ^ 1 was walked in iteration it - 1 and that iteration's barrier came after the walk: it is free, the
    // next tile's copy goes out before this tile is walked.  So do the descriptor loads of the window of the tile
    // written out in this iteration (walked L iterations ago); they are consumed after the walk.
    int ticket2 = 0;
    if (tid == 0) {
      issue_load(tile_next, s ^ 1);
      ticket2 = take();  // the tile after next (used after the walk)
    }
    const bool will_flush = depth == L;
    if (will_flush) window_load(q_tile[0]);
    const int64_t tile_start = prm.tile_origin + (int64_t)tile * kTileB;
    const int pre = (int)min((int64_t)kSandPre, tile_start);
    const int valid = (int)max((int64_t)0, min((int64_t)kTileB, prm.n - tile_start));  // text bytes of this tile
    uint8_t *txt = sh.text[s] + kSandPre;  // tile byte 0
    {
      const int staged = valid + pre;
      const int bulk = text_aligned ? (staged & ~15) : 0;
      for (int i = bulk + tid; i < staged; i += kSandThreads) txt[i - pre] = prm.text[tile_start - pre + i];
      mbar_wait(&sh.bar[s], (uint32_t)((it >> 1) & 1));
      if (bulk != staged) __syncthreads();
    }
    // ---- warm-up on the history words, then the walk
    const int chunk_rel = tid * kChunkB;
    const uint32_t Da = sand_warm<K, ADDR>(sh.B, init, txt + chunk_rel, tid == 0 ? min(warm_words, pre >> 2) : warm_words);
    const uint32_t Db = sand_warm<K, ADDR>(sh.B, init, txt + chunk_rel + kSandSplit, warm_words);
    uint32_t nz = 0;
    int cnt = 0;
    if (!(prm.debug & 2)) cnt = sand_walk<K, TAIL, ADDR>(sh.B, init, txt + chunk_rel, Da, Db, &sh.acc[0][tid], nz);
    // ---- tiles at the edges of the key window or of the text: drop hits whose start lies outside the window or
    //      whose end lies past the text (stale bytes of the buffer), bit by bit
    const bool edge = tile_start - 64 < prm.key_lo || tile_start + kTileB + 64 > prm.key_hi || valid < kTileB;
    if (edge) {
      cnt = 0;
      uint32_t nz2 = 0;
      const int64_t c0 = tile_start + chunk_rel;
      for (int a = 0; a < kAcc; ++a) {
        uint32_t m = sh.acc[a][tid], keep = 0;
        while (m) {
          const int x = sand::clz32(m);
          const uint32_t bit = 0x80000000u >> x;
          m &= ~bit;
          const int64_t end = c0 + a * G + x / K + 1;
          const int64_t start = end - (int64_t)((lens >> (8 * (x % K))) & 0xffu);
          if (end <= prm.n && start >= prm.key_lo && start < prm.key_hi) { keep |= bit; ++cnt; }
        }
        sh.acc[a][tid] = keep;
        nz2 |= min(keep, 1u) << a;
      }
      nz = nz2;
    }
    if (prm.debug & 1) { cnt = 0; nz = 0; }
    // ---- ONE barrier per tile: warp totals of the hit counts and of the flushed tile's window go to shared memory
    //      (two generations), every thread sums the eight of each it needs; the barrier also votes on ring space
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    const int lane_first = inc - cnt;                         // my first match within the warp's list
    const int warp_total = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 31) sh.scan_ws[s][warp] = inc;
    if (will_flush) {
      const unsigned long long wsum = window_sum(q_tile[0]);
      if (lane == 0) sh.red_ws[s][warp] = wsum;
    }
    if (tid == 0) sh.next_tile[s] = ticket2;  // (slot s held this iteration's tile: everyone read it before the last barrier)
    const int free_after = R - (int)(ring_tail - ring_head) + (will_flush ? q_wtotal[0] : 0);
    const int crowded = __syncthreads_or(warp_total > free_after);
    const int tile_after_next = sh.next_tile[s];
    int wfirst = 0, total = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int c = sh.scan_ws[s][w];
      if (w < warp) wfirst += c;
      total += c;
    }
    if (tid == 0) st_volatile_u64(prm.desc + tile, kFlagAgg | (unsigned long long)total);
    if (will_flush) {
      unsigned long long base = cta_base;
#pragma unroll
      for (int w = 0; w < W; ++w) base += sh.red_ws[s][w];
      flush_oldest(base);
    }
    bool direct = false;
    if (crowded) {
      // (rare) some warp's ring cannot take this tile's tokens: write out everything pending; a tile that does not
      // fit an empty ring is written out directly, in rounds, with the ring as scratch
      while (depth > 0) flush_oldest_blocking();
      direct = __syncthreads_or(warp_total > R) != 0;
    }
    uint32_t m = 0;
    int a = 0;
    if (!direct) {
      // ---- tokens of this tile -> the warp's ring, in position order (a lane's slice starts at lane_first)
      for (int k = 0; k < cnt; ++k) {
        if (m == 0) {
          a = __ffs((int)nz) - 1;
          nz &= nz - 1;
          m = sh.acc[a][tid];
        }
        int b, slot;
        sand::pop_hit<K>(m, b, slot);
        ring[(ring_tail + (uint32_t)(lane_first + k)) & (R - 1)] = (uint16_t)(((lane * kChunkB + a * G + b) << 2) | slot);
      }
      ring_tail += (uint32_t)warp_total;
#pragma unroll
      for (int i = 0; i < L; ++i)  // (depth < L here: will_flush took one out whenever it was L)
        if (i == depth) { q_tile[i] = tile; q_total[i] = total; q_wfirst[i] = wfirst; q_wtotal[i] = warp_total; }
      ++depth;
    } else {
      window_load(tile);
      const unsigned long long base = block_base(window_sum(tile));
      cta_base = base + (unsigned long long)total;
      covered = tile + 1;
      if (tid == 0 && tile == prm.n_tiles - 1) *prm.total = cta_base;
      int idx = lane_first;
      const int my_end = lane_first + cnt;
      for (int r0 = 0; r0 < warp_total; r0 += R) {
        const int lim = min(my_end, r0 + R);
        while (idx < lim) {
          if (m == 0) {
            a = __ffs((int)nz) - 1;
            nz &= nz - 1;
            m = sh.acc[a][tid];
          }
          int b, slot;
          sand::pop_hit<K>(m, b, slot);
          ring[idx - r0] = (uint16_t)(((lane * kChunkB + a * G + b) << 2) | slot);
          ++idx;
        }
        __syncwarp();
        emit(base + (unsigned long long)(wfirst + r0), warp_text_base(tile), 0u, min(R, warp_total - r0));
        __syncwarp();
      }
      ring_head = ring_tail = 0;
    }
    tile = tile_next;
    tile_next = tile_after_next;
  }
  while (depth > 0) flush_oldest_blocking();
}

template <int K, bool TAIL, int ADDR>
static void launch_sand_t(const SandParams &p, int grid, cudaStream_t st) {
  static int resident_dev[kLaunchDevSlots] = {};
  int &resident = resident_dev[launch_dev()];
  const int smem = (int)sizeof(SandShared<K>);
  if (!resident) {
    cudaFuncSetAttribute(k_sand<K, TAIL, ADDR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sand<K, TAIL, ADDR>, kSandThreads, smem);
    resident = sms * (per_sm > 0 ? per_sm : 1);
  }
  // persistent: every CTA must be resident (flushes wait for other CTAs' aggregates)
  int g = grid < resident ? grid : resident;
  if (const char *e = getenv("MRX_SAND_GRID")) g = g < atoi(e) ? g : atoi(e);  // measurement: fewer CTAs
  k_sand<K, TAIL, ADDR><<<g, kSandThreads, smem, st>>>(p);
}

template <int K>
static void launch_sand_k(const SandParams &p, bool tail, int grid, cudaStream_t st) {
  // MRX_SAND_ADDR = 0 / 1 / 2: how the walk forms the table offset (sand_entry); all are held equal by the tests,
  // the default is the one that measured fastest
  const char *e = getenv("MRX_SAND_ADDR");
  const int addr = e ? atoi(e) : 2;
  if (tail) {
    if (addr == 0) launch_sand_t<K, true, 0>(p, grid, st);
    else if (addr == 1) launch_sand_t<K, true, 1>(p, grid, st);
    else launch_sand_t<K, true, 2>(p, grid, st);
  } else {
    if (addr == 0) launch_sand_t<K, false, 0>(p, grid, st);
    else if (addr == 1) launch_sand_t<K, false, 1>(p, grid, st);
    else launch_sand_t<K, false, 2>(p, grid, st);
  }
}

void launch_sand(const SandParams &p, int k, bool tail, int grid, cudaStream_t st) {
  if (k == 1) launch_sand_k<1>(p, tail, grid, st);
  else if (k == 2) launch_sand_k<2>(p, tail, grid, st);
  else launch_sand_k<4>(p, tail, grid, st);
}

}  // namespace mrx
