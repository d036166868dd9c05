/* libmojoregex_cuda.so -- C ABI of the B200-native matching hot path of mojo-regex.
 *
 * The reference (msaelices/mojo-regex v0.21.0, pure Mojo) has no FFI seam; this ABI is
 * the seam a maintainer adds at the places where HybridMatcher calls its DFAMatcher
 * (src/regex/matcher.mojo:748-750, 793-802, 859-860, 889) and in _sub_impl_with_repl /
 * split (matcher.mojo:1679-1854, 1357-1393).  Each entry point names the reference
 * interface it replaces.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions inherited from the reference:
 *  - compiling may fail (Mojo `raises`): MRX_E_PARSE / MRX_E_UNSUPPORTED;
 *  - matching never fails: "no match" is a normal result, not an error;
 *  - spans are byte offsets into the text, end-exclusive (matching.mojo:5-46);
 *  - results borrow nothing: span arrays / output bytes are owned by the library until
 *    the matching *_free call; the text must stay alive only for the duration of a call;
 *  - one calling thread per library instance (matcher.mojo:1206-1210).
 *
 * There is no CPU fallback: every matching entry point runs CUDA kernels (sm_100a) and
 * returns MRX_E_CUDA / MRX_E_NO_DEVICE when it cannot.
 */
#ifndef MOJOREGEX_CUDA_H
#define MOJOREGEX_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRX_ABI_VERSION 1

/* return codes */
#define MRX_OK 0
#define MRX_E_INVALID (-1)     /* bad argument / corrupt blob                                   */
#define MRX_E_PARSE (-2)       /* pattern rejected by the parser (parser.mojo:225-237,340,399)   */
#define MRX_E_UNSUPPORTED (-3) /* pattern the reference routes to its NFA engines (or a path whose
                                  reference behaviour is host-SIMD-width dependent); the Mojo
                                  caller keeps its CPU engine for these, as it does today when
                                  DFA compilation raises (matcher.mojo:665-675)                  */
#define MRX_E_CUDA (-4)        /* CUDA runtime error; see mrx_last_error()                       */
#define MRX_E_NO_DEVICE (-5)   /* no CUDA device / mrx_init not called                           */
#define MRX_E_UNDEFINED (-6)   /* input on which the reference's own behaviour is undefined
                                  (split with overlapping separator spans)  */

/* where result arrays live */
#define MRX_OUT_HOST 0   /* pinned host memory, valid on return                                  */
#define MRX_OUT_DEVICE 1 /* device memory on the library stream (for callers that keep data in HBM) */

typedef struct mrx_regex mrx_regex; /* CompiledRegex: tables + dispatch plan (matcher.mojo:929-1035) */
typedef struct mrx_text mrx_text;   /* a text resident in HBM                                     */

/* MatchList replacement (matching.mojo:49-157): n spans, start_end[2*i], start_end[2*i+1]. */
typedef struct {
  int64_t *start_end;
  size_t n;
  int location; /* MRX_OUT_HOST or MRX_OUT_DEVICE */
  void *_owner;
} mrx_spans;

/* String result of sub() (matcher.mojo:1857-1917). */
typedef struct {
  uint8_t *bytes;
  size_t n;
  int location;
  void *_owner;
} mrx_bytes;

/* List[String] result of split() (matcher.mojo:1357-1393): piece i is
 * bytes[offsets[i] .. offsets[i+1]); n pieces, n+1 offsets. */
typedef struct {
  uint8_t *bytes;
  int64_t *offsets;
  size_t n;
  int location;
  void *_owner;
} mrx_pieces;

/* Introspection for tests / the Mojo side's get_stats() (matcher.mojo:1139-1163). */
typedef struct {
  int32_t n_states, n_classes, start_state;
  int32_t has_start_anchor, has_end_anchor, is_pure_literal, has_matcher, scan_eligible;
  int32_t required_byte;     /* -1 = none (matcher.mojo:558-564)                                  */
  int32_t is_wildcard_any, is_exact_literal, has_prefilter;
  int32_t findall_mode;      /* MRX_MODE_* chosen for match_all                                   */
  int32_t search_mode;       /* MRX_MODE_* chosen for match_next / sub                            */
  int32_t findall_direct;    /* 1 = potential matches are provably non-overlapping (no selection) */
  int32_t search_direct;
  int32_t fixed_sub_total_width, fixed_sub_num_groups, fixed_sub_concat;
  int32_t findall_dense, search_dense; /* 1 = served by the dense literal-set automaton (k_dense)       */
  int32_t search_rev_states;           /* > 0: batch search runs the reverse-automaton kernel (k_lines)  */
  int32_t search_skip_byte;            /* >= 0: a byte every match contains, read off the automaton; the    */
                                       /* batch kernel does not walk documents without it (-1 = none)       */
  int32_t findall_sand;                /* 1 = findall runs the Shift-And literal-set kernel (k_sand)         */
  int32_t split_fused;                 /* 1 = split runs fused (k_split): the pattern has the shape X* c Y*  */
  int32_t search_via_prefilter;        /* 1 = match_next jumps to the MemchrPrefilter literal first          */
  int32_t sub_fused_width;             /* > 0: every match is a run of exactly this many bytes of one class: */
                                       /* sub runs fused (k_subst) for replacements of up to 64 bytes         */
  int32_t findall_leading_run;         /* 1 = the pattern begins with X+ and never reads X again: match_all   */
                                       /* attempts only the first byte of every run of X (dfa.mojo:2097-2128) */
  int32_t search_dominated;            /* 1 = match_next skips candidates whose predecessor is a candidate:   */
                                       /* their attempt fails whenever the predecessor's does                 */
} mrx_info;

#define MRX_MODE_WILDCARD_ANY 0  /* `.*`                         matcher.mojo:809-813            */
#define MRX_MODE_ANCHORED 1      /* ^ or $: one attempt chain    dfa.mojo:2046-2050, 1888-1892    */
#define MRX_MODE_EXACT_LITERAL 2 /* overlapping literal hits     matcher.mojo:815-847            */
#define MRX_MODE_LITERAL 3       /* pure literal, non-overlapping dfa.mojo:2062-2071             */
#define MRX_MODE_CLASS_RUNS 4    /* maximal class runs           dfa.mojo:2075-2094              */
#define MRX_MODE_FIRST_CLASS 5   /* first-class skip + DFA       dfa.mojo:2097-2115, 2241-2253    */
#define MRX_MODE_EVERY_POS 6     /* DFA at every position        dfa.mojo:2117-2128, 1898-1903    */
#define MRX_MODE_REQUIRED_BYTE 7 /* memchr + back-scan + DFA     matcher.mojo:864-898            */
#define MRX_MODE_UNIFORM_RUN 8   /* FIRST_CLASS/EVERY_POS specialised: chain DFA over one class    */

/* ---- library ------------------------------------------------------------------------- */
int mrx_abi_version(void);
/* Select CUDA device `device` for this process and create the library stream. */
int mrx_init(int device);
/* Multi-device mode ("mrx_init(n_gpus) creates per-GPU streams"): initialise devices 0 .. n_gpus-1 in THIS
 * process.  The host-buffer entry points (mrx_findall_host, mrx_search_batch_host, mrx_sub_host, mrx_split_host)
 * then shard a large input over all of them inside one call -- byte ranges with a halo, document ranges, ranges that
 * end after a byte no match can contain -- run the shards from one worker thread per device and combine counts and
 * boundary records on the host; the result is bit for bit the single-device one.  Everything else (device-resident
 * texts, mrx_stream, profiling) addresses device 0.  The reference is single-process and single-threaded
 * (matcher.mojo:1206-1210): this is how its one caller thread reaches every GPU of the box. */
int mrx_init_multi(int n_gpus);
/* Devices in use: n_gpus after mrx_init_multi, 1 after mrx_init, 0 before either. */
int mrx_device_count(void);
int mrx_shutdown(void);
const char *mrx_last_error(void);
/* cudaStream_t the kernels are launched on (so callers can bracket them with events). */
void *mrx_stream(void);
/* Number of kernels this library has launched since mrx_init (bench.py's gpu_launches). */
uint64_t mrx_launch_count(void);
int mrx_synchronize(void);
/* Measurement aid: when enabled, the kernel a matching call is dominated by (k_dense / k_scan / k_runs /
 * k_lines; for sub and split the whole kernel sequence from match list to assembled output) is bracketed
 * by CUDA events on the library stream; mrx_last_kernel_ms() returns its duration. */
int mrx_set_profiling(int enabled);
double mrx_last_kernel_ms(void);
/* The times of the last n profiled calls, oldest first (one synchronisation; for loops of asynchronous calls).
 * Returns how many were written (at most 63). */
int mrx_kernel_ms_history(double *out_ms, int n);

/* ---- front end ------------------------------------------------------------------------ */
/* Replaces compile_regex()/CompiledRegex(pattern) (matcher.mojo:964-978, 1292-1315) for
 * callers without the Mojo front end: parse, classify, build the reference's automaton,
 * pack it.  Pure host code; needs no GPU. */
int mrx_compile(const uint8_t *pattern, size_t n, mrx_regex **out);
/* Serializer entry for the Mojo front end: pack a DFAEngine as it sits in memory
 * (dfa.mojo:257-294: `states[i].transitions` = 256 x int32 with -1 = none,
 * `states[i].is_accepting`) plus the HybridMatcher/CompiledRegex flags into a blob.
 * `matcher_lookup` = _simd_char_matcher.lookup_table (256 bytes) or NULL when
 * !_has_simd_matcher; `literal` = literal_pattern.  Call with blob == NULL to get the size. */
typedef struct {
  int32_t n_states, start_state;
  int32_t has_start_anchor, has_end_anchor, is_pure_literal, scan_eligible;
  int32_t required_byte;       /* HybridMatcher._required_byte                                  */
  int32_t is_wildcard_any, is_exact_literal, has_anchors_info;
  int32_t fixed_sub_total_width, fixed_sub_num_groups, fixed_sub_concat;
  int32_t fixed_sub_offsets[10], fixed_sub_widths[10];
} mrx_dfa_desc;
int mrx_dfa_pack(const mrx_dfa_desc *desc, const int32_t *transitions /* n_states*256 */,
                 const uint8_t *accepting /* n_states */, const uint8_t *matcher_lookup /* 256|NULL */,
                 const uint8_t *literal, size_t literal_len, const uint8_t *exact_literal,
                 size_t exact_len, const uint8_t *prefilter, size_t prefilter_len, uint8_t *blob,
                 size_t *blob_len);
/* Load a blob produced by mrx_dfa_pack / mrx_regex_blob. */
int mrx_regex_load(const uint8_t *blob, size_t n, mrx_regex **out);
int mrx_regex_blob(const mrx_regex *re, uint8_t *buf, size_t *len);
int mrx_regex_info(const mrx_regex *re, mrx_info *out);
void mrx_regex_free(mrx_regex *re);

/* ---- texts ---------------------------------------------------------------------------- */
/* Copy `n` host bytes into HBM (staged through pinned memory, async on the library stream). */
int mrx_text_upload(const uint8_t *host, size_t n, mrx_text **out);
/* Wrap `n` bytes already in device memory (not copied, not freed).  Stream contract: the library's stream is
 * created with cudaStreamNonBlocking and is NOT ordered after the stream that produced the buffer; the caller
 * synchronises its producer stream (or makes mrx_stream() wait on an event of it) before the first matching
 * call on the wrapped text. */
int mrx_text_wrap_device(const void *dptr, size_t n, mrx_text **out);
/* Multi-GPU shards: the buffer is [left halo | core | right halo].  Matching behaves as if
 * the text were exactly the buffer, but only matches whose key (start; the needle hit for
 * required-byte patterns) lies in [core_begin, core_end) are reported, and spans are
 * rebased by `global_offset` (the buffer's position in the whole corpus). */
int mrx_text_set_window(mrx_text *t, size_t core_begin, size_t core_end, int64_t global_offset);
size_t mrx_text_len(const mrx_text *t);
void mrx_text_free(mrx_text *t);

/* ---- matching (all of these launch CUDA kernels) ----------------------------------------- */
/* CompiledRegex.match_all / findall (matcher.mojo:1078-1089, 1341-1354). */
int mrx_findall(const mrx_regex *re, const mrx_text *t, int location, mrx_spans *out);
/* The same with the result left in HBM and no host synchronisation (matcher.mojo:266-270 for a caller whose
 * pipeline continues on the GPU): the scan is queued on the library stream and the call returns; out->n is
 * (size_t)-1 until mrx_spans_wait() fetches the count.  mrx_spans_count_device() is the device address of the
 * count (uint64), valid in stream order -- what mrx_comm_allgather_device() exchanges between ranks.
 * mrx_spans_wait: MRX_OK, or 1 when the span array (sized from the match density of the previous call on this
 * pattern) was too small: out->n spans are valid, the caller repeats the call (the estimate has been corrected).
 * An asynchronous result is waited for and freed by the thread that made the call (one calling thread per handle,
 * as everywhere in this ABI); freeing it while the scan is still running waits for that scan only. */
int mrx_findall_async(const mrx_regex *re, const mrx_text *t, mrx_spans *out);
const void *mrx_spans_count_device(const mrx_spans *s);
int mrx_spans_wait(mrx_spans *s);
/* findall(pattern, text) with the text in HOST memory (matcher.mojo:1341-1354: the reference's callers
 * hold the text on the host and get a host MatchList back).  Equivalent to mrx_text_upload +
 * mrx_findall(MRX_OUT_HOST) + mrx_text_free, but for large texts the upload, the scan and the download
 * of the spans are pipelined chunk by chunk over three CUDA streams.  `host` should be page-locked
 * (cudaHostAlloc / cudaHostRegister) for the copies to overlap; pageable memory works, serially. */
int mrx_findall_host(const mrx_regex *re, const uint8_t *host, size_t n, mrx_spans *out);
/* The same for one shard of a multi-GPU corpus: the host buffer is [left halo | core | right halo] with
 * the window semantics of mrx_text_set_window. */
int mrx_findall_host_shard(const mrx_regex *re, const uint8_t *host, size_t n, size_t core_begin,
                           size_t core_end, int64_t global_offset, mrx_spans *out);
/* CompiledRegex.match_next(text, start) / search (matcher.mojo:1064-1076, 1325-1338).
 * Returns 1 and fills out[0..1] on a match, 0 on no match, <0 on error. */
int mrx_search(const mrx_regex *re, const mrx_text *t, int64_t start, int64_t out[2]);
/* CompiledRegex.match_first(text, start) (matcher.mojo:1048-1061); same return convention.
 * The module-level match_first additionally requires start_idx == 0 (matcher.mojo:1411-1415);
 * that filter is the caller's. */
int mrx_match_first(const mrx_regex *re, const mrx_text *t, int64_t start, int64_t out[2]);
/* Batch search: one match_next(doc, 0) per document (the reference has no batch API; its
 * callers loop).  Document i is t[doc_offsets[i] .. doc_offsets[i+1]);
 * out_start_end[2i..2i+1] are relative to the document, (-1,-1) = None.  `doc_offsets`
 * (n_docs+1 entries) and `out_start_end` are host pointers when location == MRX_OUT_HOST,
 * device pointers when MRX_OUT_DEVICE (the results are then ordered on the library stream: use
 * mrx_stream() / mrx_synchronize() before reading them from another stream or the host).
 * doc_offsets == NULL means fixed-size documents of `doc_stride` bytes. */
int mrx_search_batch(const mrx_regex *re, const mrx_text *t, const int64_t *doc_offsets,
                     size_t n_docs, int64_t doc_stride, int location, int64_t *out_start_end);
/* sub(pattern, repl, text, count) (matcher.mojo:1857-1917, 1679-1854). count == 0: all. */
int mrx_sub(const mrx_regex *re, const uint8_t *repl, size_t repl_n, int64_t count,
            const mrx_text *t, int location, mrx_bytes *out);
/* split(pattern, text, maxsplit) (matcher.mojo:1357-1393). */
int mrx_split(const mrx_regex *re, const mrx_text *t, int64_t maxsplit, int location,
              mrx_pieces *out);

/* sub / split with the text in HOST memory (what the reference's callers pass).  Equivalent to
 * mrx_text_upload + mrx_sub / mrx_split (MRX_OUT_HOST) + mrx_text_free; large texts are cut after bytes that
 * cannot be part of any match and the chunks are uploaded, processed and downloaded as a three-stream
 * pipeline (count != 0 / maxsplit != 0, or a pattern without such bytes, take the plain path).  `host`
 * should be page-locked for the copies to overlap. */
int mrx_sub_host(const mrx_regex *re, const uint8_t *repl, size_t repl_n, int64_t count, const uint8_t *host,
                 size_t n, mrx_bytes *out);
int mrx_split_host(const mrx_regex *re, const uint8_t *host, size_t n, int64_t maxsplit, mrx_pieces *out);
/* Batch search with the text, the offsets and the results in HOST memory: groups of documents are uploaded,
 * searched and their results downloaded as a three-stream pipeline (documents are independent). */
int mrx_search_batch_host(const mrx_regex *re, const uint8_t *host, size_t n, const int64_t *doc_offsets,
                          size_t n_docs, int64_t doc_stride, int64_t *out_start_end);

/* ---- multi-GPU count exchange (one process per GPU) ---------------------------------------------------
 * The only collective on the path: an all-gather of a small int64 record per rank (match count, output
 * bytes), from which each rank derives the global offset of its slice of the result.  NCCL is looked up at
 * run time (libnccl.so.2), not linked.  Rank 0 creates the id, the caller distributes its 128 bytes (any
 * transport), every rank calls mrx_comm_init.  mrx_comm_allgather is asynchronous on the library's exchange
 * stream: `*table_dev` (world x width int64, device memory owned by the library, overwritten by the next
 * exchange) is valid after mrx_comm_join(stream) for work ordered on `stream`. */
int mrx_comm_unique_id(uint8_t out[128]);
int mrx_comm_init(const uint8_t id[128], int rank, int world);
int mrx_comm_allgather(const int64_t *values, int width /* 1..4 */, void **table_dev);
/* One int64 per rank that is still being produced on the library stream (mrx_spans_count_device): gathered
 * straight from device memory, ordered after everything queued on the library stream so far. */
int mrx_comm_allgather_device(const void *value_dev, void **table_dev);
int mrx_comm_join(void *stream /* cudaStream_t; NULL = mrx_stream() */);
int mrx_comm_shutdown(void);

void mrx_spans_free(mrx_spans *s);
void mrx_bytes_free(mrx_bytes *b);
void mrx_pieces_free(mrx_pieces *p);

#ifdef __cplusplus
}
#endif
#endif /* MOJOREGEX_CUDA_H */
