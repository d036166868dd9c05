// k_split -- split() for patterns of the shape X* c Y* (BASELINE config 5: `\s*,\s*`) as ONE pass over the text:
// separators are found, the bytes outside them are compacted into the output and the piece offsets are written
// without a match list ever reaching HBM.  (Round 1 ran k_scan -> 16 B per separator -> k_seg_lengths -> k_assemble:
// 5.8 GB of span traffic per GiB of CSV against 2.7 GB of algorithmic bytes, and 32 issue slots per comma in the
// finder; profiles/r01_k_scan_c5_ncu.txt.)  The arithmetic is in sepbits.hpp; reference semantics:
// HybridMatcher._match_all_required_byte (matcher.mojo:864-898) + split (matcher.mojo:1357-1393).
//
// A CTA owns a 36 864-byte tile, thread t its bytes [144 t, 144 t + 144):
//   1. the tile with 96 bytes before and 64 after it is staged by one TMA bulk copy;
//   2. classification: T[byte] (1 KiB table in shared memory; the lookup address depends on the text only) holds the
//      three class bits of a byte in three different bytes of a word, and acc = 2 * acc + T[byte] over the eight bytes
//      of a group, taken last to first, leaves the group's bits of each class in one byte of acc in text order -- one
//      IMAD per byte.  PRMT gathers them into 16-bit pieces per 16 text bytes, parked in three bit streams in shared
//      memory so that neighbours can read each other's bits;
//   3. every thread assembles the 288-bit window around its chunk and runs sepb::separators (multi-word additions
//      whose carries follow the runs): kept-byte mask, valid-separator mask, overlap and overflow flags;
//   4. (separators, kept bytes) are scanned across the block as one packed integer and published as the tile's
//      aggregate; the descriptor loads of the look-back window go out here and are consumed after step 5;
//   5. the kept bytes are compacted IN PLACE: every thread takes its 144 text bytes into registers, and after a
//      barrier writes the kept ones back to the same buffer from its rank on, four bytes per store (PRMT with a
//      selector looked up by the word's four keep bits packs them; a 64-bit queue carries the remainder);  every
//      valid separator parks the number of kept bytes before it (tile relative, 16 bits) in a second staging array;
//   6. with the tile's base known (sliding window over the aggregates, as in k_sand) both staging arrays are written
//      out with coalesced stores: the bytes 16 at a time, the piece offsets as base + staged value.
// The class table is replicated 32 times (entry b of lane l at word 32 b + l): every lane reads its own bank, so the
// lookups are conflict free whatever the text is (the first version shared one 1 KiB table: upper- and lower-case
// letters and digits collide there, 148 M of its 266 M shared-memory wavefronts per GiB were bank conflicts,
// profiles/r02_k_split_a_ncu.txt).
// Tiles are handed out by ticket; persistent CTAs, every CTA resident (a flush waits for other CTAs' aggregates).
#include "kernels.cuh"
#include "sepbits.hpp"

namespace mrx {

constexpr int kSplitPre = 96;    // staged bytes before the tile: the window's history
constexpr int kSplitPost = 64;   // ... and after it (48 are used)
constexpr int kSplitHw = kThreads * kVecs;  // 16-bit pieces of one tile's bit stream

constexpr int kSplitOffCap = 8192;  // separators of one tile whose offsets are staged (more: written directly)

struct SplitShared {
  alignas(128) uint8_t text[kSplitPre + kTile + kSplitPost];  // the staged text; then, in place, the tile's kept bytes
  uint32_t T[256 * 32];                        // class table, one copy per lane: entry b of lane l at [32 b + l]
  uint16_t bits[3][6 + kSplitHw + 4];          // C / L / bad streams: 6 pieces of history, the tile, 3 of look-ahead
  uint16_t offs[kSplitOffCap];                 // kept bytes of the tile before each of its separators
  uint32_t sel[16];                            // PRMT selectors that pack the kept bytes of a word (by its 4 keep bits)
  alignas(8) uint64_t bar;
  unsigned long long red_ws[2][kThreads / 32];
  uint32_t scan_ws[kThreads / 32];
  int next_tile;
};

// class bits of 16 text bytes: three 16-bit pieces (bit i <-> byte i)
// T = this lane's column of the replicated table (byte offset 4 * lane already added)
__device__ __forceinline__ void split_classify16(const uint32_t *T, const uint4 &v, uint32_t &c, uint32_t &l, uint32_t &b) {
  auto entry = [&](uint32_t w, int k) -> uint32_t {
    const uint32_t x = __byte_perm(w, 0, 0x4440 + k);
    uint32_t off;
    asm("mad.lo.u32 %0, %1, 128, 0;" : "=r"(off) : "r"(x));
    return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(T) + off);
  };
  auto group = [&](uint32_t w0, uint32_t w1) -> uint32_t {  // 8 bytes, last first: byte i ends in bit i of each class byte
    uint32_t acc = 0;
#pragma unroll
    for (int k = 3; k >= 0; --k) acc = acc * 2u + entry(w1, k);
#pragma unroll
    for (int k = 3; k >= 0; --k) acc = acc * 2u + entry(w0, k);
    return acc;
  };
  const uint32_t lo = group(v.x, v.y), hi = group(v.z, v.w);
  c = __byte_perm(lo, hi, 0x3340);  // (byte 3 of an acc is zero)
  l = __byte_perm(lo, hi, 0x3351);
  b = __byte_perm(lo, hi, 0x3362);
}

__global__ void __launch_bounds__(kThreads, 2) k_split(const SplitParams prm) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SplitShared &sh = *reinterpret_cast<SplitShared *>(smem_raw);
  constexpr int W = kThreads / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool text_aligned = (reinterpret_cast<uintptr_t>(prm.text) & 15) == 0;
  if (tid == 0) {
    mbar_init(&sh.bar, 1);
    fence_proxy_async();
  }
  for (int i = tid; i < 256 * 32; i += kThreads) sh.T[i] = prm.table->T[i >> 5];
  if (tid < 16) {  // selector that moves the kept bytes of a word to its low end, in order (the rest: byte 0 of the 2nd operand)
    uint32_t sel = 0x4444u;
    int o = 0;
    for (int i = 0; i < 4; ++i)
      if ((tid >> i) & 1) { sel = (sel & ~(0xfu << (4 * o))) | ((uint32_t)i << (4 * o)); ++o; }
    sh.sel[tid] = sel;
  }
  const uint32_t *myT = sh.T + lane;
  __syncthreads();

  auto issue_load = [&](int t) {  // thread 0 only
    if (t >= prm.n_tiles) return;
    const int64_t ts = (int64_t)t * kTile;
    const int pre = (int)min((int64_t)kSplitPre, ts);
    int64_t bytes = min((int64_t)kTile + kSplitPost, prm.n - ts) + pre;
    if (bytes < 0) bytes = 0;
    const uint32_t bulk = text_aligned ? ((uint32_t)bytes & ~15u) : 0u;
    mbar_expect_tx(&sh.bar, bulk);
    if (bulk) tma_bulk_g2s(sh.text + (kSplitPre - pre), prm.text + ts - pre, bulk, &sh.bar);
  };

  // ordered output: exclusive prefixes (kept bytes, separators) of everything before tile `covered`
  unsigned long long base_k = 0, base_s = 0;
  int covered = 0;
  unsigned long long win[2];
  auto window_load = [&](int p) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = p - 1 - (tid + kThreads * u);
      win[u] = i >= covered ? ld_volatile_u64(prm.desc + i) : kFlagAgg;
    }
  };
  auto window_sum = [&](int p, unsigned long long &sk, unsigned long long &ss) {  // this warp's part, both counters
    sk = ss = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = p - 1 - (tid + kThreads * u);
      while ((win[u] >> 62) == 0) win[u] = ld_volatile_u64(prm.desc + i);
      sk += win[u] & 0xfffffull;
      ss += (win[u] >> 20) & 0xfffffull;
    }
    for (int i = p - 1 - 2 * kThreads - tid; i >= covered; i -= kThreads) {  // (rare) the rest of a long window
      unsigned long long d;
      do { d = ld_volatile_u64(prm.desc + i); } while ((d >> 62) == 0);
      sk += d & 0xfffffull;
      ss += (d >> 20) & 0xfffffull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sk += __shfl_xor_sync(0xffffffffu, sk, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
  };

  if (tid == 0) {
    sh.next_tile = (int)atomicAdd(prm.ticket, 1u);
    issue_load(sh.next_tile);
  }
  __syncthreads();
  int tile = sh.next_tile;
  for (int it = 0; tile < prm.n_tiles; ++it) {
    int ticket = 0;
    if (tid == 0) ticket = (int)atomicAdd(prm.ticket, 1u);  // the next tile; used after the compaction
    const int64_t tile_start = (int64_t)tile * kTile;
    const int pre = (int)min((int64_t)kSplitPre, tile_start);
    const int64_t left = prm.n - tile_start;                                   // text bytes from the tile start on
    const int valid = (int)min((int64_t)kTile, left);                          // ... inside the tile
    const int staged = (int)min((int64_t)kTile + kSplitPost, left) + pre;      // bytes in the buffer from (tile_start - pre)
    uint8_t *txt = sh.text + kSplitPre;                                        // tile byte 0
    {
      const int bulk = text_aligned ? (staged & ~15) : 0;
      for (int i = bulk + tid; i < staged; i += kThreads) txt[i - pre] = prm.text[tile_start - pre + i];
      mbar_wait(&sh.bar, (uint32_t)(it & 1));
      if (bulk != staged) __syncthreads();
    }
    // ---- 2. class bits of my chunk -> own words + the three streams
    const int chunk_rel = tid * kChunk;
    const int cv = max(0, min(kChunk, valid - chunk_rel));  // bytes of my chunk that exist
    uint32_t oc[5], ol[5], ob[5];                           // own bits: words 0..3 full, word 4 half
#pragma unroll
    for (int j = 0; j < 5; ++j) oc[j] = ol[j] = ob[j] = 0;
#pragma unroll
    for (int v = 0; v < kVecs; ++v) {
      const uint4 tv = *reinterpret_cast<const uint4 *>(txt + chunk_rel + 16 * v);
      uint32_t c, l, b;
      split_classify16(myT, tv, c, l, b);
      const int live = cv - 16 * v;
      if (live < 16) {
        const uint32_t m = live <= 0 ? 0u : ((1u << live) - 1u);
        c &= m; l &= m; b &= m;
      }
      const int h = 6 + kVecs * tid + v;
      sh.bits[0][h] = (uint16_t)c; sh.bits[1][h] = (uint16_t)l; sh.bits[2][h] = (uint16_t)b;
      oc[v >> 1] |= c << (16 * (v & 1)); ol[v >> 1] |= l << (16 * (v & 1)); ob[v >> 1] |= b << (16 * (v & 1));
    }
    if (tid < 9) {  // the 6 pieces before the tile and the 3 after it, one 16-byte group each
      const int rel = tid < 6 ? -kSplitPre + 16 * tid : kTile + 16 * (tid - 6);  // first byte, relative to the tile
      const uint4 tv = *reinterpret_cast<const uint4 *>(txt + rel);
      uint32_t c, l, b;
      split_classify16(myT, tv, c, l, b);
      // bytes that exist: [-pre, left)
      uint32_t m = 0xffffu;
      if (rel < -pre) m = (rel + 16 <= -pre) ? 0u : (m & ~((1u << (-pre - rel)) - 1u));
      const int64_t live = left - rel;
      if (live < 16) m &= live <= 0 ? 0u : ((1u << (int)live) - 1u);
      const int h = tid < 6 ? tid : 6 + kSplitHw + (tid - 6);
      sh.bits[0][h] = (uint16_t)(c & m); sh.bits[1][h] = (uint16_t)(l & m); sh.bits[2][h] = (uint16_t)(b & m);
    }
    __syncthreads();  // #1
    // ---- 3. the 288-bit window: 6 pieces of the left neighbour, my 9, 3 of the right neighbour
    sepb::Win C, L, B;
    {
      const uint16_t *sc = sh.bits[0] + kVecs * tid, *sl = sh.bits[1] + kVecs * tid, *sb = sh.bits[2] + kVecs * tid;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        C.w[k] = (uint32_t)sc[2 * k] | ((uint32_t)sc[2 * k + 1] << 16);
        L.w[k] = (uint32_t)sl[2 * k] | ((uint32_t)sl[2 * k + 1] << 16);
        B.w[k] = (uint32_t)sb[2 * k] | ((uint32_t)sb[2 * k + 1] << 16);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { C.w[3 + j] = oc[j]; L.w[3 + j] = ol[j]; B.w[3 + j] = ob[j]; }
      C.w[7] = oc[4] | ((uint32_t)sc[15] << 16);
      L.w[7] = ol[4] | ((uint32_t)sl[15] << 16);
      B.w[7] = ob[4] | ((uint32_t)sb[15] << 16);
      C.w[8] = (uint32_t)sc[16] | ((uint32_t)sc[17] << 16);
      L.w[8] = (uint32_t)sl[16] | ((uint32_t)sl[17] << 16);
      B.w[8] = (uint32_t)sb[16] | ((uint32_t)sb[17] << 16);
    }
    sepb::Own own = sepb::separators(C, L, B);
    {
      // a run of L bytes the window cannot measure: two adjacent 16-byte groups entirely in L (own groups 0..8 and
      // the first one after the chunk)
      bool ovf = false;
#pragma unroll
      for (int j = 3; j <= 7; ++j) {
        ovf = ovf || L.w[j] == 0xffffffffu;
        if (j < 7) ovf = ovf || __funnelshift_r(L.w[j], L.w[j + 1], 16) == 0xffffffffu;
      }
      int fl = (ovf ? 1 : 0) | (own.overlap ? 2 : 0);
      if (fl) atomicOr(prm.flags, fl);
    }
    uint32_t nk = 0, ns = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      // bytes past the end of the text are neither kept nor separators
      const int live = cv - 32 * j;
      const uint32_t m = live >= 32 ? 0xffffffffu : (live <= 0 ? 0u : ((1u << live) - 1u));
      own.kept[j] &= m;
      nk += __popc(own.kept[j]);
      ns += __popc(own.sep[j]);
    }
    // ---- 4. one packed scan: separators in the high half, kept bytes in the low half (both < 2^16 per tile)
    uint32_t pk = (ns << 16) | nk, inc = pk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) sh.scan_ws[warp] = inc;
    __syncthreads();  // #2
    uint32_t before = inc - pk, total = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const uint32_t c = sh.scan_ws[w];
      if (w < warp) before += c;
      total += c;
    }
    const uint32_t rank_k = before & 0xffffu, rank_s = before >> 16;
    const uint32_t total_k = total & 0xffffu, total_s = total >> 16;
    if (tid == 0) st_volatile_u64(prm.desc + tile, kFlagAgg | ((unsigned long long)total_s << 20) | total_k);
    window_load(tile);  // consumed after the compaction
    // ---- 5a. my text -> registers; the barrier below frees the buffer for the compacted bytes
    uint4 mine[kVecs];
#pragma unroll
    for (int v = 0; v < kVecs; ++v) mine[v] = *reinterpret_cast<const uint4 *>(txt + chunk_rel + 16 * v);
    // ---- 5b. piece offsets of my separators (tile relative) -> staging array, or straight to HBM later
    const bool offs_staged = total_s <= (uint32_t)kSplitOffCap;
    if (offs_staged) {
      uint32_t below = rank_k, idx = rank_s;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        uint32_t m = own.sep[j];
        while (m) {
          const int b = __ffs((int)m) - 1;
          m &= m - 1;
          sh.offs[idx++] = (uint16_t)(below + __popc(own.kept[j] & ((1u << b) - 1u)));
        }
        below += __popc(own.kept[j]);
      }
    }
    if (tid == 0) sh.next_tile = ticket;
    __syncthreads();  // #3: every thread holds its text; staged offsets and the next ticket are posted
    // ---- 5c. kept bytes -> the text buffer, in place, from my rank on.  A 64-bit queue (lo, hi) collects packed
    //      bytes; full words are stored aligned.  The first word of a thread is shared with its predecessor and the
    //      last one with its successor: those two are written byte by byte.
    uint8_t *stage = sh.text;
    {
      uint32_t lo = 0, hi = 0;
      uint32_t fill = rank_k & 3u;                 // bytes of the current word that are not mine (first word only)
      uint8_t *wp = stage + (rank_k & ~3u);        // aligned address of the current word
      bool first = true;
      const uint32_t skip = fill;
#pragma unroll
      for (int v = 0; v < kVecs; ++v) {
        const uint32_t ws[4] = {mine[v].x, mine[v].y, mine[v].z, mine[v].w};
        const uint32_t k16 = own.kept[v >> 1] >> (16 * (v & 1));
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
          const uint32_t nib = (k16 >> (4 * wi)) & 15u;
          const uint32_t packed = __byte_perm(ws[wi], 0, sh.sel[nib]);
          const uint32_t s8 = fill * 8u;
          lo |= packed << s8;
          hi |= __funnelshift_l(packed, 0u, s8);   // packed >> (32 - s8), 0 for s8 == 0
          fill += __popc(nib);
          if (fill >= 4u) {
            if (first) {
              for (uint32_t q = skip; q < 4u; ++q) wp[q] = (uint8_t)(lo >> (8 * q));
              first = false;
            } else {
              *reinterpret_cast<uint32_t *>(wp) = lo;
            }
            wp += 4;
            lo = hi;
            hi = 0;
            fill -= 4u;
          }
        }
      }
      for (uint32_t q = first ? skip : 0u; q < fill; ++q) wp[q] = (uint8_t)(lo >> (8 * q));  // the last, partial word
    }
    // ---- the look-back window, loaded before the compaction: by now its tiles have been counted
    unsigned long long sk, ss;
    window_sum(tile, sk, ss);
    if (lane == 0) { sh.red_ws[0][warp] = sk; sh.red_ws[1][warp] = ss; }
    __syncthreads();  // #4: the compacted tile is complete, the window sums are posted
    unsigned long long bk = base_k, bs = base_s;
#pragma unroll
    for (int w = 0; w < W; ++w) { bk += sh.red_ws[0][w]; bs += sh.red_ws[1][w]; }
    base_k = bk + total_k;
    base_s = bs + total_s;
    covered = tile + 1;
    if (tid == 0 && tile == prm.n_tiles - 1) { prm.totals[0] = base_k; prm.totals[1] = base_s; }
    // ---- 6b. piece offsets: the k-th valid separator starts piece k + 1 at (kept bytes before it)
    if (offs_staged) {
      for (uint32_t i = tid; i < total_s; i += kThreads) {
        const unsigned long long idx = bs + i;
        if (idx < (unsigned long long)prm.offs_cap) prm.offs[1 + idx] = (int64_t)(bk + sh.offs[i]);
      }
    } else {  // (a tile with more than kSplitOffCap separators) every thread writes its own
      unsigned long long idx = bs + rank_s;
      const unsigned long long at = bk + rank_k;
      uint32_t below = 0;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        uint32_t m = own.sep[j];
        while (m) {
          const int b = __ffs((int)m) - 1;
          m &= m - 1;
          const uint32_t in_word = __popc(own.kept[j] & ((1u << b) - 1u));
          if (idx < (unsigned long long)prm.offs_cap) prm.offs[1 + idx] = (int64_t)(at + below + in_word);
          ++idx;
        }
        below += __popc(own.kept[j]);
      }
    }
    // ---- 6a. the tile's kept bytes -> out[bk, bk + total_k): head up to 16-byte alignment, 16-byte body, tail
    {
      uint8_t *dst = prm.out + bk;
      const int nkept = (int)total_k;
      int head = (int)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
      if (head > nkept) head = nkept;
      if (tid < head) dst[tid] = stage[tid];
      const int body = (nkept - head) >> 4;
      const int sh8 = (head & 3) * 8;
      const uint32_t *sw = reinterpret_cast<const uint32_t *>(stage + (head & ~3));
      for (int i = tid; i < body; i += kThreads) {
        const uint32_t a0 = sw[4 * i], a1 = sw[4 * i + 1], a2 = sw[4 * i + 2], a3 = sw[4 * i + 3], a4 = sw[4 * i + 4];
        uint4 o;
        o.x = __funnelshift_r(a0, a1, sh8); o.y = __funnelshift_r(a1, a2, sh8);
        o.z = __funnelshift_r(a2, a3, sh8); o.w = __funnelshift_r(a3, a4, sh8);
        *reinterpret_cast<uint4 *>(dst + head + 16 * i) = o;
      }
      const int done = head + 16 * body;
      if (tid < nkept - done) dst[done + tid] = stage[done + tid];
    }
    __syncthreads();  // #5: the buffer is free for the next tile; everyone has read sh.next_tile's predecessor
    if (tid == 0) issue_load(sh.next_tile);
    tile = sh.next_tile;
  }
}

void launch_split(const SplitParams &p, cudaStream_t st) {
  static int resident_dev[kLaunchDevSlots] = {};
  int &resident = resident_dev[launch_dev()];
  const int smem = (int)sizeof(SplitShared);
  if (!resident) {
    cudaFuncSetAttribute(k_split, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_split, kThreads, smem);
    resident = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int grid = p.n_tiles < resident ? p.n_tiles : resident;
  if (grid > 0) k_split<<<grid, kThreads, smem, st>>>(p);
}

}  // namespace mrx
