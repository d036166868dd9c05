// Kernel parameter blocks and host-side launchers (definitions in kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mojoregex_cuda.h"
#include "device.cuh"
#include "program.hpp"

namespace mrx {

enum OutKind : int32_t {
  OUT_SPANS = 0,  // ordered (start,end) list through look-back compaction
  OUT_FIRST = 1,  // global minimum key (search)
  OUT_BATCH = 2,  // per-document minimum (batch search)
};

struct ScanParams {
  const uint8_t *text;
  int64_t n;            // bytes in the buffer; the reference sees text[0, n)
  int64_t tile_origin;  // first tile starts here (16-byte aligned position)
  int64_t key_lo, key_hi;  // candidate positions reported: [key_lo, key_hi)
  int32_t key_hi_inclusive_n;  // position n itself is a candidate (empty match at end of text)
  int32_t try_at_end;
  int64_t search_start;
  const DevProg *prog;
  Filter filt;
  int32_t mode;
  int32_t out_kind;
  int32_t dominated;
  int32_t n_tiles;
  int64_t *spans;
  int64_t *keys;
  int64_t cap;
  int64_t rebase;
  unsigned long long *desc;
  unsigned int *ticket;
  unsigned long long *total;
  unsigned long long *best;
  const int64_t *doc_offsets;
  int64_t doc_stride;
  int64_t n_docs;
  unsigned long long *doc_best;
  const unsigned long long *doc_min;  // OUT_BATCH behind a MemchrPrefilter: (first literal occurrence of the document,
                                      // relative) << 32 | its end, ~0 = none; attempts before it are not made
};

struct RunsParams {
  const uint8_t *text;
  int64_t n;
  int64_t tile_origin;
  int64_t key_lo, key_hi;
  const DevProg *prog;
  Filter filt;
  int32_t mode;  // MRX_MODE_CLASS_RUNS or MRX_MODE_UNIFORM_RUN
  int32_t run_min, run_max;
  int32_t n_tiles;
  int32_t tile_base;    // tiles [tile_base, n_tiles) are handed out by this launch (ticket starts at 0)
  const long long *tile_carry;
  int64_t *spans;
  int64_t cap;
  int64_t rebase;
  unsigned long long *desc;
  unsigned int *ticket;
  unsigned long long *total;
};

// Dense literal-set automaton in HBM: header, then uint16 table[n_states * 256] (entry = next * 512).
struct DenseHdr {
  int32_t n_states, acc_base, max_len, table_bytes;
  int32_t uniform_len, pad0, pad1, pad2;  // > 0 when every literal has this length
  uint8_t out_len[64];
  uint8_t len_lut[32];  // literal length by (row offset >> 3) & 31 of the accept row (rows are 520 B apart,
                        // so consecutive rows differ by 1 there)
};

struct DenseParams {
  const uint8_t *text;
  int64_t n;
  int64_t tile_origin;
  int64_t key_lo, key_hi;  // match starts reported: [key_lo, key_hi)
  const uint8_t *dense;    // DenseHdr + table
  int32_t table_bytes;
  int32_t out_kind;        // OUT_SPANS or OUT_FIRST
  int32_t n_tiles;
  int32_t tile_base;       // tiles [tile_base, n_tiles) are handed out by this launch (ticket starts at 0)
  int64_t *spans;
  int64_t cap;
  int64_t rebase;
  unsigned long long *desc;
  unsigned int *ticket;
  unsigned long long *total;
  unsigned long long *best;  // OUT_FIRST: min over (start << 8 | len)
};

// Shift-And literal-set scan (k_sand, sand.hpp); its tiles are kSandTileBytes long (256 threads x 144 bytes)
namespace sand { struct Table; }
constexpr int kSandTileBytes = 36864;
struct SandParams {
  const uint8_t *text;
  int64_t n;
  int64_t tile_origin;
  int64_t key_lo, key_hi;      // match starts reported: [key_lo, key_hi)
  const sand::Table *table;    // in HBM
  int32_t n_tiles;
  int32_t tile_base;           // this launch scans tiles [tile_base, n_tiles)
  unsigned long long base0;    // matches before tile_base (the total of the launches that came before)
  int64_t *spans;
  int64_t cap;
  int64_t rebase;
  unsigned long long *desc;    // one word per tile, zeroed: (1 << 62) | aggregate once the tile is counted
  unsigned int *ticket;        // zeroed: tiles are handed out in order, tile = tile_base + ticket
  unsigned long long *total;
  int32_t debug;               // measurement only (MRX_SAND_DEBUG): 1 = no matches kept, 2 = no walk (results are
                               // invalid), 4 = static round-robin tiles instead of tickets
};

// Fused split for patterns of the shape X* c Y* (k_split, sepbits.hpp)
namespace sepb { struct Table; }
struct SplitParams {
  const uint8_t *text;
  int64_t n;
  const sepb::Table *table;    // in HBM
  int32_t n_tiles;             // ceil(n / kTile)
  uint8_t *out;                // the kept bytes (at most n, + 16 slack)
  int64_t *offs;               // offs[1 + k] = kept bytes before the k-th separator; offs[0] and the last entry: host
  int64_t offs_cap;            // separators that can be recorded
  unsigned long long *desc;    // one word per tile, zeroed: (1 << 62) | separators << 20 | kept bytes
  unsigned int *ticket;        // zeroed
  unsigned long long *totals;  // [0] kept bytes, [1] separators of the whole text
  int *flags;                  // zeroed; bit 0: a run the window cannot measure, bit 1: overlapping separators
};

// Fused sub for fixed-width chain automata over one class (k_subst)
constexpr int kSubstMaxRepl = 64;     // bytes of one expanded replacement
constexpr int kSubstMaxStage = 72 * 1024;
struct SubstParams {
  const uint8_t *text;
  int64_t n;
  const DevProg *prog;         // cand[] = the class
  Filter filt;
  int32_t width;               // W: bytes per match, 4..32
  int32_t repl_len;            // R: bytes per expanded replacement, <= kSubstMaxRepl
  const int32_t *tpl;          // R entries: >= 0 byte of the match at this offset, < 0 literal byte -1 - value
  int32_t n_tiles;             // ceil(n / kTile)
  int32_t stage_bytes;         // shared-memory staging: >= kTile + max(0, R - W) * (kTile / W + 1) + 96
  const long long *tile_carry; // in-class bytes immediately before every tile (k_tile_trailing + k_carry_scan)
  uint8_t *out;                // n + max(0, R - W) * (n / W) + 16 bytes
  unsigned long long *desc;    // one word per tile, zeroed: (1 << 62) | matches that start in the tile
  unsigned int *ticket;        // zeroed
  unsigned long long *totals;  // [0] output bytes, [1] matches
};

// Per-document search (k_lines): one thread per document.
struct LinesParams {
  const uint8_t *text;
  int64_t n;
  const uint8_t *tables;       // RevTable::rev followed by RevTable::fwd, in HBM (16-byte aligned sizes)
  int32_t rev_bytes, fwd_bytes;
  int32_t fwd_start;           // row offset of the forward start state
  const int64_t *doc_offsets;  // n_docs + 1 entries, or nullptr for fixed-size documents
  int64_t doc_stride;
  int64_t n_docs;
  int64_t *out;                // (start, end) relative to the document, (-1, -1) = no match
  int32_t staged;              // 1: fixed-size, 16-byte aligned documents staged through shared memory by TMA
  int32_t line_pitch;          // shared-memory bytes per staged document (odd multiple of 16)
  int32_t req_byte;            // RevTable::req_byte: documents without this byte cannot match (-1: none)
  unsigned int *ticket;        // k_lines_req: zeroed counter, batches of 32 documents are handed out by it
};

enum SegMode : int32_t { SEG_SUB = 0, SEG_SPLIT = 1 };

constexpr int kMaxTplSegs = 32;
struct SubTemplate {
  int32_t n_segs;
  int32_t group_off[kMaxTplSegs];  // >= 0: copy `len` bytes from match_start + group_off; < 0: literal
  int32_t lit_off[kMaxTplSegs];
  int32_t len[kMaxTplSegs];
  const uint8_t *literal;  // replacement string in device memory
};

// sub / split output assembly (k_assemble): one CTA per kAsmTile output bytes
constexpr int kAsmRounds = 2;
constexpr int kAsmTile = kThreads * 16 * kAsmRounds;  // 8 KiB
constexpr int kAsmSegCap = 1536;                       // segment records staged per tile (24 KiB)
constexpr int kAsmTplCap = 256;
constexpr int kSegRounds = 8;                          // k_seg_lengths: segments per thread
constexpr int kSegTile = kThreads * kSegRounds;        // ... and per CTA                        // replacement bytes staged per tile
struct AsmParams {
  const uint8_t *text;
  int64_t n;
  const int64_t *spans;
  int64_t M;
  const int64_t *offs;      // M + 2 output offsets of the segments (k_seg_lengths)
  const int64_t *tile_seg;  // segment owning the first byte of every output tile
  int64_t n_tiles;
  int32_t mode;             // SegMode
  int64_t repl_len;
  const int32_t *tpl_src;   // flattened replacement, repl_len entries
  int64_t total;            // output bytes
  uint8_t *out;             // 16-byte aligned, at least total + 16 bytes
};

// Function attributes (opt-in dynamic shared memory) and occupancy are per DEVICE: the launchers keep what they have
// set up in arrays indexed by the current device (multi-device mode launches from one worker thread per device, each
// touching only its own slot).
constexpr int kLaunchDevSlots = 16;
inline int launch_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d & (kLaunchDevSlots - 1);
}

void launch_scan(const ScanParams &p, int grid, cudaStream_t st);
void launch_lines(LinesParams p, cudaStream_t st);
void launch_runs(const RunsParams &p, int grid, cudaStream_t st);
void launch_dense(const DenseParams &p, int grid, cudaStream_t st);
void launch_sand(const SandParams &p, int k, bool tail, int grid, cudaStream_t st);
void launch_split(const SplitParams &p, cudaStream_t st);
void launch_subst(const SubstParams &p, cudaStream_t st);
void launch_tile_trailing(const uint8_t *text, int64_t n, int64_t origin, int n_tiles, const DevProg *prog,
                          long long *trail, int *all, cudaStream_t st);
void launch_carry_scan(int n_tiles, const long long *trail, const int *all, long long *carry, const long long *carry0,
                       cudaStream_t st);
void launch_back_run(const uint8_t *text, int64_t origin, const DevProg *prog, long long *out, cudaStream_t st);
void launch_attempt_one(const ScanParams &p, int64_t pos, int64_t *out, cudaStream_t st);
void launch_batch_trivial(const ScanParams &p, int wildcard, int64_t *out, cudaStream_t st);
void launch_batch_finalize(const unsigned long long *doc_best, int64_t n_docs, const int64_t *doc_offsets,
                           int64_t doc_stride, int64_t n, int empty_at_end, int64_t *out, cudaStream_t st);
void launch_fill_u64(unsigned long long *p, int64_t n, unsigned long long v, cudaStream_t st);
void launch_add_i64(int64_t *p, int64_t n, int64_t v, cudaStream_t st);
void launch_select_prefix_max(const int64_t *spans, int64_t m, unsigned long long *desc, unsigned int *ticket,
                              int64_t *pmax, cudaStream_t st);
void launch_select_walk(const int64_t *spans, const int64_t *keys, const int64_t *pmax, int64_t m, uint8_t *sel,
                        cudaStream_t st);
void launch_select_prefilter(const int64_t *spans, int64_t m, const int64_t *lits, int64_t l, uint8_t *sel, cudaStream_t st);
void launch_select_compact(const int64_t *spans, const uint8_t *sel, int64_t m, unsigned long long *desc,
                           unsigned int *ticket, int64_t *out, unsigned long long *total, cudaStream_t st);
void launch_seg_lengths(const int64_t *spans, int64_t M, int64_t n, int mode, int64_t repl_len,
                        unsigned long long *desc, unsigned int *ticket, int64_t *offs, int *bad, int64_t *tile_seg,
                        cudaStream_t st);
void launch_assemble(const AsmParams &p, cudaStream_t st);

}  // namespace mrx
