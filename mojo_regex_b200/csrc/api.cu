// C ABI of libmojoregex_cuda.so (see include/mojoregex_cuda.h).  Host orchestration only:
// every byte of text is touched by the kernels in kernels.cu, never by host code.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "frontend.hpp"
#include "kernels.cuh"
#include "program.hpp"

using namespace mrx;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

struct Lib {
  bool inited = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the host-buffer pipelines (H2D / D2H)
  std::vector<cudaEvent_t> ev_pool;              // untimed events for those pipelines
  uint64_t launches = 0;
  // caching allocators: device blocks and pinned host blocks, bucketed by power-of-two size
  std::multimap<size_t, void *> dev_free, host_free;
  std::map<void *, size_t> dev_live, host_live;
  // small persistent scratch
  unsigned int *ticket = nullptr;        // 4 tickets
  unsigned long long *scalars = nullptr; // [0] total, [1] best, [2..4] attempt_one out
  unsigned long long *h_scalars = nullptr;  // pinned mirror
  bool profiling = false;
  bool prof_outer = false;  // sub / split bracket their whole kernel sequence; list jobs then do not record
  // kernel timing for the bench: a ring of event pairs, so that a loop of asynchronous calls can be timed launch by
  // launch without a synchronisation in between (mrx_kernel_ms_history)
  static constexpr int kProfRing = 64;
  cudaEvent_t evs[kProfRing][2] = {};
  unsigned ev_idx = 0, ev_count = 0;
  cudaEvent_t begin_ev() { ++ev_idx; if (ev_count < kProfRing - 1) ++ev_count; return evs[ev_idx % kProfRing][0]; }
  cudaEvent_t end_ev() { return evs[ev_idx % kProfRing][1]; }
  bool ev_valid = false;
};
// One Lib per CUDA device.  Single-device mode (mrx_init) uses the slot of the selected device; multi-device mode
// (mrx_init_multi) initialises slots 0 .. n-1 and the host-buffer entry points shard their input over them, one
// worker thread per device.  `g` is the calling thread's current device.
constexpr int kMaxDevices = 16;
Lib g_devs[kMaxDevices];
int g_multi = 0;                       // > 1: devices 0 .. g_multi - 1 are initialised
thread_local Lib *g_cur = &g_devs[0];
#define g (*g_cur)

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) return fail(MRX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// Pool bucket sizes: powers of two up to 64 MiB, then eighths of the enclosing power of two (at most
// 12.5 % over-allocation for the multi-gigabyte text, span and output buffers).
size_t bucket(size_t n) {
  size_t b = 256;
  while (b < n) b <<= 1;
  if (b <= (64u << 20)) return b;
  const size_t step = b >> 4;  // b/2 < n <= b: steps of (b/2)/8
  return (n + step - 1) / step * step;
}
int dev_alloc(size_t n, void **out) {
  const size_t b = bucket(n ? n : 1);
  auto it = g.dev_free.find(b);
  if (it != g.dev_free.end()) {
    *out = it->second;
    g.dev_free.erase(it);
  } else {
    cudaError_t e = cudaMalloc(out, b);
    if (e != cudaSuccess) {
      // drop the cache and retry once
      for (auto &kv : g.dev_free) cudaFree(kv.second);
      g.dev_free.clear();
      e = cudaMalloc(out, b);
      if (e != cudaSuccess) return fail(MRX_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
  }
  g.dev_live[*out] = b;
  return MRX_OK;
}
void dev_release(void *p) {
  if (!p) return;
  auto it = g.dev_live.find(p);
  if (it == g.dev_live.end()) return;
  g.dev_free.emplace(it->second, p);
  g.dev_live.erase(it);
}
int host_alloc(size_t n, void **out) {
  const size_t b = bucket(n ? n : 1);
  auto it = g.host_free.find(b);
  if (it != g.host_free.end()) {
    *out = it->second;
    g.host_free.erase(it);
  } else {
    cudaError_t e = cudaHostAlloc(out, b, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(MRX_E_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  }
  g.host_live[*out] = b;
  return MRX_OK;
}
void host_release(void *p) {
  if (!p) return;
  auto it = g.host_live.find(p);
  if (it == g.host_live.end()) return;
  g.host_free.emplace(it->second, p);
  g.host_live.erase(it);
}

// RAII for temporaries taken from the device pool
struct DevTmp {
  void *p = nullptr;
  ~DevTmp() { dev_release(p); }
  int get(size_t n) { return dev_alloc(n, &p); }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// a pool block that is handed on with take() or goes back to the pool when the scope is left early
struct PoolBlock {
  void *p = nullptr;
  ~PoolBlock() { dev_release(p); }
  void *take() { void *q = p; p = nullptr; return q; }
};
// sub / split bracket their whole kernel sequence: list jobs must not record while the flag is up, and the flag
// must come down on every exit path
struct ProfOuter {
  bool armed = false;
  void arm() { if (g.profiling) { g.prof_outer = armed = true; cudaEventRecord(g.begin_ev(), g.stream); } }
  void disarm() { if (armed) { g.prof_outer = false; armed = false; } }
  ~ProfOuter() { disarm(); }
};
// result owner that is released again unless the call commits it (early CUDA_TRY returns included)
struct OwnerGuard {
  void *own;
  void **slot;
  ~OwnerGuard();
  void commit() { own = nullptr; }
};
// the host-buffer pipelines own device / pinned blocks that async copies on the side streams still use: drain
// those streams before the blocks can go back to the pool, whatever the exit path
struct DrainSideStreams {
  ~DrainSideStreams() { if (g.s_in) cudaStreamSynchronize(g.s_in); if (g.s_out) cudaStreamSynchronize(g.s_out); }
};

}  // namespace

// the tables of a regex as they sit in one device's HBM, and what that device learned about the texts
struct RegexOnDevice {
  DevProg *d_findall = nullptr, *d_search = nullptr, *d_prefilter = nullptr;
  uint8_t *d_dense_findall = nullptr, *d_dense_search = nullptr;  // DenseHdr + table
  sand::Table *d_sand_findall = nullptr;                          // Shift-And table of the findall plan (k_sand)
  sepb::Table *d_sep_findall = nullptr;                           // class table of the fused split (k_split)
  uint8_t *d_rev_search = nullptr;                                // RevTable::tbl of the search plan
  // matches per key byte seen by the last list job of each plan (findall, search): sizes the span
  // buffer of the next call so that the scan does not have to be repeated with a larger one
  double density[2] = {0.0, 0.0};
};
struct mrx_regex {
  Program prog;
  bool search_dominated = false;
  bool findall_leading_run = false;  // compute_leading_run(findall plan)
  RegexOnDevice dev[16];  // by device index (kMaxDevices); a worker thread only touches its own device's slot
  RegexOnDevice &on() { return dev[g_cur->device < 0 ? 0 : g_cur->device]; }
  const RegexOnDevice &on() const { return dev[g_cur->device < 0 ? 0 : g_cur->device]; }
};

struct mrx_text {
  uint8_t *d = nullptr;
  size_t n = 0;
  bool owned = false;
  size_t core_begin = 0, core_end = 0;
  int64_t global_offset = 0;
};

namespace {

struct OwnedOut {  // behind the _owner fields
  void *dev[3] = {nullptr, nullptr, nullptr};
  void *host[3] = {nullptr, nullptr, nullptr};
  Lib *lib = g_cur;  // the blocks belong to this device's pools
};
void release_owner(void *o) {
  if (!o) return;
  OwnedOut *w = static_cast<OwnedOut *>(o);
  Lib *saved = g_cur;
  g_cur = w->lib;
  for (void *p : w->dev) dev_release(p);
  for (void *p : w->host) host_release(p);
  g_cur = saved;
  delete w;
}
OwnerGuard::~OwnerGuard() {
  if (own) { release_owner(own); *slot = nullptr; }
}

// MRX_DEBUG_SYNC=1: synchronise after the named launch and report which one failed (debugging aid)
int debug_sync(const char *what) {
  static const bool on = getenv("MRX_DEBUG_SYNC") != nullptr;
  if (!on) return MRX_OK;
  cudaError_t e = cudaStreamSynchronize(g.stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MRX_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return MRX_OK;
}

int require_device() {
  if (!g.inited) return fail(MRX_E_NO_DEVICE, "mrx_init() has not been called (no CUDA device selected); there is no CPU fallback");
  return MRX_OK;
}

Filter make_filter(const Plan &pl) {
  Filter f{};
  f.kind = pl.filt;
  auto splat = [](uint32_t b) { return b * 0x01010101u; };
  if (pl.filt == FILT_NEEDLES) {
    f.n_needles = pl.n_needles;
    f.n0 = splat(pl.needles[0]); f.n1 = splat(pl.needles[1]); f.n2 = splat(pl.needles[2]); f.n3 = splat(pl.needles[3]);
  } else if (pl.filt == FILT_RANGE) {
    f.n0 = splat(pl.range_lo);
    f.hi_is_ff = pl.range_hi == 255;
    f.n1 = splat((uint32_t)((pl.range_hi + 1) & 255));
  }
  return f;
}

// a table that lives as long as its regex: plain cudaMalloc (not the pool), freed again if the copy fails
int upload_blob(const void *src, size_t n, void **out) {
  void *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, n));
  cudaError_t e = cudaMemcpyAsync(d, src, n, cudaMemcpyHostToDevice, g.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(MRX_E_CUDA, std::string("table upload: ") + cudaGetErrorString(e));
  }
  *out = d;
  return MRX_OK;
}

int upload_devprog(const Program &p, const Plan &pl, DevProg **out) {
  // attempts walk a dense forward table (one lookup per byte) when it fits the table budget, else the
  // class-compressed one (byte -> class, (state, class) -> state, state -> accepting)
  std::vector<uint16_t> fwd;
  uint16_t fwd_start = 0;
  if (p.n_states > 0) build_forward_table(p, fwd, fwd_start);
  const bool dense_fwd = p.n_states > 0 && fwd.size() * 2 <= (size_t)kMaxTrans;
  if (!dense_fwd && (size_t)p.n_states * p.n_classes > (size_t)kMaxTrans)
    return fail(MRX_E_UNSUPPORTED, "automaton too large for the shared-memory table (" + std::to_string(p.n_states) + " states x " +
                                       std::to_string(p.n_classes) + " classes)");
  if (pl.literal.size() > (size_t)kMaxLiteral) return fail(MRX_E_UNSUPPORTED, "literal longer than 256 bytes");
  std::vector<uint8_t> raw(sizeof(DevProg), 0);
  DevProg *h = reinterpret_cast<DevProg *>(raw.data());
  h->n_states = p.n_states;
  h->n_classes = p.n_classes;
  h->start_state = p.start_state;
  h->flags = (p.flag(BF_END_ANCHOR) ? DF_END_ANCHOR : 0) | (p.flag(BF_START_ANCHOR) ? DF_START_ANCHOR : 0) |
             (p.start_accepting() ? DF_START_ACCEPTING : 0) | (pl.shortcut ? DF_SHORTCUT : 0) |
             (dense_fwd ? DF_DENSE_FWD : 0);
  h->fwd_start = fwd_start;
  h->literal_len = (int32_t)pl.literal.size();
  h->required_byte = p.required_byte;
  h->trans_len = dense_fwd ? (int32_t)(fwd.size() * 2) : p.n_states * p.n_classes;
  std::memcpy(h->byte_class, p.byte_class.data(), 256);
  std::memcpy(h->lookup, p.lookup.data(), 256);
  std::memcpy(h->cand, pl.cand.data(), 256);
  if (!p.accepting.empty()) std::memcpy(h->accepting, p.accepting.data(), p.accepting.size());
  if (!pl.literal.empty()) std::memcpy(h->literal, pl.literal.data(), pl.literal.size());
  if (dense_fwd) std::memcpy(h->trans, fwd.data(), fwd.size() * 2);
  else if (!p.trans.empty()) std::memcpy(h->trans, p.trans.data(), p.trans.size());
  void *d = nullptr;
  if (int rc = upload_blob(h, sizeof(DevProg), &d)) return rc;
  *out = static_cast<DevProg *>(d);
  return MRX_OK;
}

int upload_dense(const DenseTable &dt, uint8_t **out) {
  std::vector<uint8_t> raw(sizeof(DenseHdr) + dt.next.size() * 2, 0);
  DenseHdr *h = reinterpret_cast<DenseHdr *>(raw.data());
  h->n_states = dt.n_states; h->acc_base = dt.acc_base; h->max_len = dt.max_len;
  h->table_bytes = (int32_t)(dt.next.size() * 2);
  h->uniform_len = 0;
  if (!dt.out_len.empty()) {
    bool same = true;
    for (uint8_t l : dt.out_len) same = same && l == dt.out_len[0];
    if (same) h->uniform_len = dt.out_len[0];
  }
  if (dt.out_len.size() > sizeof(h->out_len)) return fail(MRX_E_UNSUPPORTED, "too many literals for the dense table");
  std::memcpy(h->out_len, dt.out_len.data(), dt.out_len.size());
  if (dt.out_len.size() > 32) return fail(MRX_E_UNSUPPORTED, "too many literals for the dense table");
  std::memcpy(h->len_lut, dt.len_lut.data(), 32);
  std::memcpy(raw.data() + sizeof(DenseHdr), dt.next.data(), dt.next.size() * 2);
  void *d = nullptr;
  if (int rc = upload_blob(raw.data(), raw.size(), &d)) return rc;
  *out = static_cast<uint8_t *>(d);
  return MRX_OK;
}

int ensure_uploaded(mrx_regex *re) {
  if (re->prog.flag(BF_WILDCARD_ANY)) return MRX_OK;
  if (re->prog.findall.dense.valid && !re->on().d_dense_findall) { int rc = upload_dense(re->prog.findall.dense, &re->on().d_dense_findall); if (rc) return rc; }
  if (re->prog.search.dense.valid && !re->on().d_dense_search) { int rc = upload_dense(re->prog.search.dense, &re->on().d_dense_search); if (rc) return rc; }
  if (re->prog.findall.sep_valid && !re->on().d_sep_findall) {
    void *d = nullptr;
    if (int rc = upload_blob(&re->prog.findall.sep, sizeof(sepb::Table), &d)) return rc;
    re->on().d_sep_findall = static_cast<sepb::Table *>(d);
  }
  if (re->prog.findall.sand_valid && !re->on().d_sand_findall) {
    void *d = nullptr;
    if (int rc = upload_blob(&re->prog.findall.sand, sizeof(sand::Table), &d)) return rc;
    re->on().d_sand_findall = static_cast<sand::Table *>(d);
  }
  if (re->prog.search.rev.valid && !re->on().d_rev_search) {
    const RevTable &rt = re->prog.search.rev;
    std::vector<uint16_t> tb(rt.rev);
    tb.insert(tb.end(), rt.fwd.begin(), rt.fwd.end());
    void *d = nullptr;
    if (int rc = upload_blob(tb.data(), tb.size() * 2, &d)) return rc;
    re->on().d_rev_search = static_cast<uint8_t *>(d);
  }
  if (!re->on().d_findall) { int rc = upload_devprog(re->prog, re->prog.findall, &re->on().d_findall); if (rc) return rc; }
  if (!re->on().d_search && !re->prog.search.unsupported) { int rc = upload_devprog(re->prog, re->prog.search, &re->on().d_search); if (rc) return rc; }
  if (re->prog.search_via_prefilter && !re->on().d_prefilter) { int rc = upload_devprog(re->prog, re->prog.prefilter_plan, &re->on().d_prefilter); if (rc) return rc; }
  return MRX_OK;
}

int read_scalars(int first, int count) {
  CUDA_TRY(cudaMemcpyAsync(g.h_scalars + first, g.scalars + first, sizeof(unsigned long long) * count, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return MRX_OK;
}

// A(p) failing implies A(p+1) failing when both bytes are candidates: after two bytes the younger
// attempt sits in the state the older one is in ("dominated candidates").
bool compute_dominated(const Program &p, const Plan &pl) {
  if (pl.shortcut || p.start_accepting() || p.n_states == 0) return false;
  if (pl.mode != MRX_MODE_FIRST_CLASS && pl.mode != MRX_MODE_EVERY_POS && pl.mode != MRX_MODE_UNIFORM_RUN) return false;
  const int nc = p.n_classes, s0 = p.start_state;
  std::vector<uint8_t> ccls(nc, 0);
  for (int b = 0; b < 256; ++b) if (pl.cand[b]) ccls[p.byte_class[b]] = 1;
  // candidates must be exactly class-aligned for the argument below
  for (int b = 0; b < 256; ++b) if ((pl.cand[b] != 0) != (ccls[p.byte_class[b]] != 0)) return false;
  for (int a = 0; a < nc; ++a) {
    if (!ccls[a]) continue;
    int sa = p.trans[(size_t)s0 * nc + a];
    for (int b = 0; b < nc; ++b) {
      if (!ccls[b]) continue;
      int via = sa == kDead ? kDead : p.trans[(size_t)sa * nc + b];
      int direct = p.trans[(size_t)s0 * nc + b];
      if (via != direct) return false;
      // the younger attempt must not accept where the older one did not: state after `b` is shared,
      // and the older one's state after `a` alone is irrelevant to the younger.
    }
  }
  return true;
}

// match_all (dfa.mojo:2097-2128) makes an attempt at every candidate that no earlier match covers.  For patterns that
// begin with X+ and never read a byte of X again (`[a-z]+[0-9]+`, `a+b`) only the first byte of a run of X can start a
// match that is reported.  Condition on the automaton: R = the states reached from the start state by bytes of X
// alone; all of R agree on acceptance and on every byte outside X (so how many bytes of X an attempt has read does
// not matter), and every state outside R and the start state is dead on X.  Then the attempt at the first byte of a
// run walks the whole run inside R: it either covers the run or fails, and every later attempt in the run, being in
// R at the end of the run too, fails the same way; no match ends inside a run it did not start in (it would have to
// read a byte of X outside its leading run).  Without this a text that is one long run costs run^2 / 2 walked bytes.
// (`a+1a?` is not in the family: "a1aa1" -> (0,3), (3,5), the second match starts on an `a` that follows an `a`.)
bool compute_leading_run(const Program &p, const Plan &pl) {
  if (pl.shortcut || p.start_accepting() || p.n_states == 0) return false;
  if (pl.mode != MRX_MODE_FIRST_CLASS && pl.mode != MRX_MODE_EVERY_POS) return false;
  const int nc = p.n_classes, s0 = p.start_state;
  std::vector<uint8_t> ccls(nc, 0);
  for (int b = 0; b < 256; ++b) if (pl.cand[b]) ccls[p.byte_class[b]] = 1;
  for (int b = 0; b < 256; ++b) if ((pl.cand[b] != 0) != (ccls[p.byte_class[b]] != 0)) return false;
  std::vector<uint8_t> inR(p.n_states, 0);
  std::vector<int> work{s0};
  int first = -1;
  while (!work.empty()) {
    const int s = work.back();
    work.pop_back();
    for (int a = 0; a < nc; ++a) {
      if (!ccls[a]) continue;
      const int to = p.trans[(size_t)s * nc + a];
      if (to == kDead || to == s0) return false;  // a run of X must keep the attempt alive, and not restart it
      if (!inR[to]) { inR[to] = 1; work.push_back(to); if (first < 0) first = to; }
    }
  }
  if (first < 0) return false;
  for (int s = 0; s < p.n_states; ++s) {
    if (inR[s]) {
      if (p.accepting[s] != p.accepting[first]) return false;
      for (int a = 0; a < nc; ++a)
        if (!ccls[a] && p.trans[(size_t)s * nc + a] != p.trans[(size_t)first * nc + a]) return false;
    } else if (s != s0) {
      for (int a = 0; a < nc; ++a)
        if (ccls[a] && p.trans[(size_t)s * nc + a] != kDead) return false;
    }
  }
  return true;
}

struct DevSpans {
  int64_t *spans = nullptr;  // pool block
  int64_t n = 0;
};

// ---- ordered match list for a plan --------------------------------------------------------------
struct ListJob {
  const mrx_regex *re;
  const mrx_text *t;
  const Plan *pl;
  const DevProg *dprog;
  int64_t key_lo, key_hi;  // buffer coordinates
  int64_t rebase;
  const uint8_t *dense = nullptr;  // DenseHdr + table when the plan has a dense automaton
  const sand::Table *sand = nullptr;  // ... and its Shift-And form when the set fits (findall plan only)
  // mrx_findall_async: the kernel leaves its match count here (device memory owned by the result) and the call
  // returns without reading it; *async_cap = the capacity of the span array that was launched with
  unsigned long long *async_total = nullptr;
  int64_t *async_cap = nullptr;
};

// MRX_NO_SAND=1: literal sets go through k_dense even when k_sand applies (the tests hold the two equal)
bool sand_enabled() { return getenv("MRX_NO_SAND") == nullptr; }

// Span-buffer capacity for a list job: a guess of one match per 8 bytes (at most 2^26 spans) the first
// time, then 1/8 above the density the same plan saw last time.  A scan that overflows its buffer still
// reports the exact count, and is repeated once with that.
double &density_slot(const ListJob &job) {
  mrx_regex *re = const_cast<mrx_regex *>(job.re);
  return re->on().density[job.pl == &job.re->prog.findall ? 0 : 1];
}
int64_t initial_cap(const ListJob &job) {
  const int64_t keys = job.key_hi - job.key_lo;
  const double d = density_slot(job);
  if (d > 0.0) return (int64_t)(d * 1.125 * (double)keys) + 4096;
  return std::min<int64_t>(keys / 8 + 1024, 1ll << 26);
}
void note_density(const ListJob &job, int64_t total) {
  const int64_t keys = job.key_hi - job.key_lo;
  if (keys > 0) density_slot(job) = (double)total / (double)keys;
}

int64_t n_tiles_for(int64_t origin, int64_t hi_inclusive, int64_t tile_bytes = kTile) {
  if (hi_inclusive < origin) return 0;
  return (hi_inclusive - origin) / tile_bytes + 1;
}

int run_scan_list(const ListJob &job, bool want_keys, DevSpans &out, int64_t **keys_out) {
  const Plan &pl = *job.pl;
  const int64_t n = (int64_t)job.t->n;
  const bool at_end = pl.try_at_end && job.re->prog.start_accepting() && job.key_hi >= n;
  const int64_t origin = job.key_lo & ~15ll;
  int64_t last_pos = at_end ? n : job.key_hi - 1;
  // the dense kernel records a match in the tile that holds its END byte: a match that starts inside the key
  // window may end up to max_len - 1 bytes past it (its start is filtered against the window by the kernel)
  if (job.dense && !want_keys) last_pos = std::min<int64_t>(n - 1, job.key_hi - 1 + std::max(pl.dense.max_len - 1, 0));
  const bool use_sand = job.sand && job.dense && !want_keys;
  const int64_t tiles = n_tiles_for(origin, last_pos, use_sand ? kSandTileBytes : kTile);
  out.n = 0;
  if (tiles <= 0) return MRX_OK;
  if (tiles > 0x7fffffff) return fail(MRX_E_INVALID, "text too large for one launch");
  int64_t cap = initial_cap(job);
  DevTmp desc;
  int rc = desc.get(sizeof(unsigned long long) * tiles);
  if (rc) return rc;
  for (int attempt = 0; attempt < 3; ++attempt) {
    PoolBlock b_spans, b_keys;
    if ((rc = dev_alloc(sizeof(int64_t) * 2 * cap, &b_spans.p))) return rc;
    if (want_keys && (rc = dev_alloc(sizeof(int64_t) * cap, &b_keys.p))) return rc;
    void *spans = b_spans.p, *keys = b_keys.p;
    CUDA_TRY(cudaMemsetAsync(desc.p, 0, sizeof(unsigned long long) * tiles, g.stream));
    CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
    unsigned long long *d_total = job.async_total ? job.async_total : g.scalars;
    CUDA_TRY(cudaMemsetAsync(d_total, 0, 8, g.stream));
    ScanParams p{};
    p.text = job.t->d; p.n = n; p.tile_origin = origin; p.key_lo = job.key_lo; p.key_hi = job.key_hi;
    p.key_hi_inclusive_n = at_end ? 1 : 0; p.try_at_end = at_end ? 1 : 0; p.search_start = job.key_lo;
    p.prog = job.dprog; p.filt = make_filter(pl); p.mode = pl.mode; p.out_kind = OUT_SPANS;
    p.dominated = (job.pl == &job.re->prog.findall && job.re->findall_leading_run) ? 1 : 0;
    p.n_tiles = (int)tiles; p.spans = (int64_t *)spans; p.keys = (int64_t *)keys; p.cap = cap; p.rebase = job.rebase;
    p.desc = desc.as<unsigned long long>(); p.ticket = g.ticket; p.total = d_total;
    if (g.profiling && !g.prof_outer) cudaEventRecord(g.begin_ev(), g.stream);
    if (use_sand) {
      SandParams sp{};
      sp.text = job.t->d; sp.n = n; sp.tile_origin = origin; sp.key_lo = job.key_lo; sp.key_hi = job.key_hi;
      sp.table = job.sand; sp.n_tiles = (int)tiles; sp.tile_base = 0; sp.base0 = 0; sp.spans = (int64_t *)spans; sp.cap = cap;
      sp.rebase = job.rebase; sp.desc = desc.as<unsigned long long>(); sp.ticket = g.ticket; sp.total = d_total;
      if (const char *dbg = getenv("MRX_SAND_DEBUG")) sp.debug = atoi(dbg);
      launch_sand(sp, pl.sand.k, pl.sand.tail != 0, (int)tiles, g.stream);
    } else if (job.dense && !want_keys) {
      DenseParams dp{};
      dp.text = job.t->d; dp.n = n; dp.tile_origin = origin; dp.key_lo = job.key_lo; dp.key_hi = job.key_hi;
      dp.dense = job.dense; dp.table_bytes = (int32_t)(pl.dense.next.size() * 2); dp.out_kind = OUT_SPANS;
      dp.n_tiles = (int)tiles; dp.spans = (int64_t *)spans; dp.cap = cap; dp.rebase = job.rebase;
      dp.desc = desc.as<unsigned long long>(); dp.ticket = g.ticket; dp.total = d_total;
      launch_dense(dp, (int)tiles, g.stream);
    } else {
      launch_scan(p, (int)tiles, g.stream);
    }
    if (g.profiling && !g.prof_outer) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    if (job.async_total) {  // no host round trip: the caller reads the count later (mrx_spans_wait)
      out.spans = (int64_t *)b_spans.take();
      out.n = -1;
      *job.async_cap = cap;
      return MRX_OK;
    }
    if ((rc = read_scalars(0, 1))) return rc;
    const int64_t total = (int64_t)g.h_scalars[0];
    note_density(job, total);
    if (total <= cap) {
      out.spans = (int64_t *)b_spans.take();
      out.n = total;
      if (keys_out) *keys_out = (int64_t *)b_keys.take();
      return MRX_OK;
    }
    cap = total;
  }
  return fail(MRX_E_CUDA, "match count did not stabilise");
}

int run_runs_list(const ListJob &job, DevSpans &out) {
  const Plan &pl = *job.pl;
  const int64_t n = (int64_t)job.t->n;
  const int64_t origin = job.key_lo & ~15ll;
  const int64_t tiles = n_tiles_for(origin, job.key_hi - 1);
  out.n = 0;
  if (tiles <= 0) return MRX_OK;
  if (tiles > 0x7fffffff) return fail(MRX_E_INVALID, "text too large for one launch");
  DevTmp desc, trail, allf, carry, carry0;
  int rc = desc.get(sizeof(unsigned long long) * tiles);
  if (rc) return rc;
  const long long *d_carry = nullptr;
  if (pl.mode == MRX_MODE_UNIFORM_RUN) {
    if ((rc = trail.get(sizeof(long long) * tiles))) return rc;
    if ((rc = allf.get(sizeof(int) * tiles))) return rc;
    if ((rc = carry.get(sizeof(long long) * tiles))) return rc;
    // the carry of the first tile is the length of the class run that ends right before `origin` (a windowed
    // text: the kernel sees the whole buffer, so a run straddling core_begin keeps its phase)
    const long long *d_carry0 = nullptr;
    if (origin > 0) {
      if ((rc = carry0.get(sizeof(long long)))) return rc;
      launch_back_run(job.t->d, origin, job.dprog, carry0.as<long long>(), g.stream);
      ++g.launches;
      d_carry0 = carry0.as<long long>();
    }
    launch_tile_trailing(job.t->d, n, origin, (int)tiles, job.dprog, trail.as<long long>(), allf.as<int>(), g.stream);
    launch_carry_scan((int)tiles, trail.as<long long>(), allf.as<int>(), carry.as<long long>(), d_carry0, g.stream);
    g.launches += 2;
    d_carry = carry.as<long long>();
  }
  int64_t cap = initial_cap(job);
  for (int attempt = 0; attempt < 3; ++attempt) {
    PoolBlock b_spans;
    if ((rc = dev_alloc(sizeof(int64_t) * 2 * cap, &b_spans.p))) return rc;
    void *spans = b_spans.p;
    CUDA_TRY(cudaMemsetAsync(desc.p, 0, sizeof(unsigned long long) * tiles, g.stream));
    CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
    unsigned long long *d_total = job.async_total ? job.async_total : g.scalars;
    CUDA_TRY(cudaMemsetAsync(d_total, 0, 8, g.stream));
    RunsParams p{};
    p.text = job.t->d; p.n = n; p.tile_origin = origin; p.key_lo = job.key_lo; p.key_hi = job.key_hi;
    p.prog = job.dprog; p.filt = make_filter(pl); p.mode = pl.mode; p.run_min = pl.run_min; p.run_max = pl.run_max;
    p.n_tiles = (int)tiles; p.tile_carry = d_carry; p.spans = (int64_t *)spans; p.cap = cap; p.rebase = job.rebase;
    p.desc = desc.as<unsigned long long>(); p.ticket = g.ticket; p.total = d_total;
    if (g.profiling && !g.prof_outer) cudaEventRecord(g.begin_ev(), g.stream);
    launch_runs(p, (int)tiles, g.stream);
    if (g.profiling && !g.prof_outer) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    if (job.async_total) {
      out.spans = (int64_t *)b_spans.take();
      out.n = -1;
      *job.async_cap = cap;
      return MRX_OK;
    }
    if ((rc = read_scalars(0, 1))) return rc;
    const int64_t total = (int64_t)g.h_scalars[0];
    note_density(job, total);
    if (total <= cap) { out.spans = (int64_t *)b_spans.take(); out.n = total; return MRX_OK; }
    cap = total;
  }
  return fail(MRX_E_CUDA, "match count did not stabilise");
}

int select_list(DevSpans &pot, int64_t *keys, DevSpans &out) {
  const int64_t m = pot.n;
  out.n = 0;
  if (m == 0) { out.spans = pot.spans; pot.spans = nullptr; dev_release(keys); return MRX_OK; }
  const int64_t tiles = (m + kThreads - 1) / kThreads;
  DevTmp desc, pmax, sel;
  int rc;
  if ((rc = desc.get(sizeof(unsigned long long) * tiles))) return rc;
  if ((rc = pmax.get(sizeof(int64_t) * m))) return rc;
  if ((rc = sel.get((size_t)m))) return rc;
  void *outp = nullptr;
  if ((rc = dev_alloc(sizeof(int64_t) * 2 * m, &outp))) return rc;
  CUDA_TRY(cudaMemsetAsync(desc.p, 0, sizeof(unsigned long long) * tiles, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
  CUDA_TRY(cudaMemsetAsync(sel.p, 0, (size_t)m, g.stream));
  launch_select_prefix_max(pot.spans, m, desc.as<unsigned long long>(), g.ticket, pmax.as<int64_t>(), g.stream);
  launch_select_walk(pot.spans, keys, pmax.as<int64_t>(), m, sel.as<uint8_t>(), g.stream);
  CUDA_TRY(cudaMemsetAsync(desc.p, 0, sizeof(unsigned long long) * tiles, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.scalars, 0, 8, g.stream));
  launch_select_compact(pot.spans, sel.as<uint8_t>(), m, desc.as<unsigned long long>(), g.ticket, (int64_t *)outp, g.scalars, g.stream);
  g.launches += 3;
  CUDA_TRY(cudaGetLastError());
  if ((rc = read_scalars(0, 1))) { dev_release(outp); return rc; }
  out.spans = (int64_t *)outp;
  out.n = (int64_t)g.h_scalars[0];
  dev_release(pot.spans);
  pot.spans = nullptr;
  dev_release(keys);
  return MRX_OK;
}

int search_impl(const mrx_regex *re, const mrx_text *t, int64_t start, int64_t out[2]);
thread_local bool search_skip_prefilter = false;  // set while search_impl runs the search that follows the literal jump

// Final ordered, selected match list for `pl` over the text's core window.
int match_list(const mrx_regex *re, const mrx_text *t, const Plan &pl, const DevProg *dprog, DevSpans &out) {
  out = DevSpans();
  if (pl.unsupported) return fail(MRX_E_UNSUPPORTED, pl.why);
  const int64_t n = (int64_t)t->n;
  ListJob job{re, t, &pl, dprog, (int64_t)t->core_begin, (int64_t)t->core_end, t->global_offset};
  if (pl.dense.valid) job.dense = (&pl == &re->prog.findall) ? re->on().d_dense_findall : re->on().d_dense_search;
  if (&pl == &re->prog.findall && pl.sand_valid && sand_enabled()) job.sand = re->on().d_sand_findall;
  switch (pl.mode) {
    case MRX_MODE_WILDCARD_ANY: {  // matcher.mojo:809-813
      void *p = nullptr;
      int rc = dev_alloc(16, &p);
      if (rc) return rc;
      int64_t h[2] = {t->global_offset, n + t->global_offset};
      CUDA_TRY(cudaMemcpyAsync(p, h, 16, cudaMemcpyHostToDevice, g.stream));
      CUDA_TRY(cudaStreamSynchronize(g.stream));
      out.spans = (int64_t *)p;
      out.n = 1;
      return MRX_OK;
    }
    case MRX_MODE_ANCHORED: {  // dfa.mojo:2046-2050: one match_next(text, 0)
      int64_t se[2];
      int rc = search_impl(re, t, 0, se);
      if (rc < 0) return rc;
      if (rc == 0 || se[0] < (int64_t)t->core_begin || se[0] >= std::max<int64_t>((int64_t)t->core_end, 1)) {
        if (!(rc == 1 && se[0] == n && (int64_t)t->core_end == n)) return MRX_OK;
      }
      void *p = nullptr;
      if ((rc = dev_alloc(16, &p))) return rc;
      int64_t h[2] = {se[0] + t->global_offset, se[1] + t->global_offset};
      CUDA_TRY(cudaMemcpyAsync(p, h, 16, cudaMemcpyHostToDevice, g.stream));
      CUDA_TRY(cudaStreamSynchronize(g.stream));
      out.spans = (int64_t *)p;
      out.n = 1;
      return MRX_OK;
    }
    case MRX_MODE_CLASS_RUNS:
    case MRX_MODE_UNIFORM_RUN:
      return run_runs_list(job, out);
    default: {
      if (&pl == &re->prog.search && re->prog.search_via_prefilter) {
        // every potential match of the search plan, every occurrence of the prefilter literal, then the chain
        // pos -> next literal -> next potential match (k_select_prefilter)
        DevSpans pot, lits;
        int rc = run_scan_list(job, false, pot, nullptr);
        if (rc) return rc;
        PoolBlock b_pot{pot.spans};
        ListJob lj{re, t, &re->prog.prefilter_plan, re->on().d_prefilter, job.key_lo, job.key_hi, job.rebase};
        if ((rc = run_scan_list(lj, false, lits, nullptr))) return rc;
        PoolBlock b_lits{lits.spans};
        if (pot.n == 0 || lits.n == 0) return MRX_OK;
        const int64_t m = pot.n, tiles = (m + kThreads - 1) / kThreads;
        DevTmp desc, sel;
        PoolBlock b_out;
        if ((rc = desc.get(sizeof(unsigned long long) * tiles))) return rc;
        if ((rc = sel.get((size_t)m))) return rc;
        if ((rc = dev_alloc(sizeof(int64_t) * 2 * m, &b_out.p))) return rc;
        CUDA_TRY(cudaMemsetAsync(sel.p, 0, (size_t)m, g.stream));
        CUDA_TRY(cudaMemsetAsync(desc.p, 0, sizeof(unsigned long long) * tiles, g.stream));
        CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
        CUDA_TRY(cudaMemsetAsync(g.scalars, 0, 8, g.stream));
        launch_select_prefilter(pot.spans, m, lits.spans, lits.n, sel.as<uint8_t>(), g.stream);
        launch_select_compact(pot.spans, sel.as<uint8_t>(), m, desc.as<unsigned long long>(), g.ticket, (int64_t *)b_out.p, g.scalars, g.stream);
        g.launches += 2;
        CUDA_TRY(cudaGetLastError());
        if ((rc = read_scalars(0, 1))) return rc;
        out.n = (int64_t)g.h_scalars[0];
        out.spans = (int64_t *)b_out.take();
        return MRX_OK;
      }
      if (pl.direct) return run_scan_list(job, false, out, nullptr);
      DevSpans pot;
      int64_t *keys = nullptr;
      const bool want_keys = pl.mode == MRX_MODE_REQUIRED_BYTE;
      int rc = run_scan_list(job, want_keys, pot, &keys);
      if (rc) return rc;
      return select_list(pot, keys, out);
    }
  }
}

int search_impl(const mrx_regex *re, const mrx_text *t, int64_t start, int64_t out[2]) {
  const Plan &pl = re->prog.search;
  const int64_t n = (int64_t)t->n;
  if (start < 0) start = 0;
  if (pl.mode == MRX_MODE_WILDCARD_ANY) {  // matcher.mojo:762-766
    if (start <= n) { out[0] = start; out[1] = n; return 1; }
    return 0;
  }
  if (pl.unsupported) return fail(MRX_E_UNSUPPORTED, pl.why);
  if (re->prog.search_via_prefilter && !search_skip_prefilter) {
    // matcher.mojo:784-796: c = first occurrence of the prefilter literal at or after start (none: no match), then
    // match_next from c
    if (start >= n) return 0;  // prefilter.mojo:421-422
    const Plan &pre = re->prog.prefilter_plan;
    ScanParams p{};
    p.text = t->d; p.n = n; p.prog = re->on().d_prefilter; p.filt = make_filter(pre); p.mode = pre.mode;
    p.search_start = start; p.key_lo = start; p.key_hi = n; p.out_kind = OUT_FIRST; p.best = g.scalars + 1;
    launch_fill_u64(g.scalars + 1, 1, ~0ull, g.stream);
    ++g.launches;
    const int64_t origin = start & ~15ll, tiles = n_tiles_for(origin, n - 1);
    int64_t done = 0, win = 64;
    while (done < tiles) {
      const int64_t cnt = std::min(win, tiles - done);
      p.tile_origin = origin + done * kTile;
      p.n_tiles = (int)cnt;
      launch_scan(p, (int)cnt, g.stream);
      ++g.launches;
      CUDA_TRY(cudaGetLastError());
      int rc = read_scalars(1, 1);
      if (rc) return rc;
      if (g.h_scalars[1] != ~0ull) break;
      done += cnt;
      win *= 4;
    }
    if (g.h_scalars[1] == ~0ull) return 0;
    search_skip_prefilter = true;
    const int rc = search_impl(re, t, (int64_t)g.h_scalars[1], out);
    search_skip_prefilter = false;
    return rc;
  }
  if (pl.dense.valid && re->on().d_dense_search) {
    if (pl.mode == MRX_MODE_EXACT_LITERAL && start >= n) return 0;  // matcher.mojo:773-774
    if (start >= n) return 0;
    const int64_t origin = start & ~15ll;
    const int64_t tiles = n_tiles_for(origin, n - 1);
    DenseParams dp{};
    dp.text = t->d; dp.n = n; dp.key_lo = start; dp.key_hi = n; dp.dense = re->on().d_dense_search;
    dp.table_bytes = (int32_t)(pl.dense.next.size() * 2); dp.out_kind = OUT_FIRST; dp.best = g.scalars + 1;
    launch_fill_u64(g.scalars + 1, 1, ~0ull, g.stream);
    ++g.launches;
    int64_t done = 0, win = 64;
    while (done < tiles) {
      const int64_t cnt = std::min(win, tiles - done);
      dp.tile_origin = origin + done * kTile;
      dp.n_tiles = (int)cnt;
      launch_dense(dp, (int)cnt, g.stream);
      ++g.launches;
      CUDA_TRY(cudaGetLastError());
      int rc = read_scalars(1, 1);
      if (rc) return rc;
      if (g.h_scalars[1] != ~0ull) break;
      done += cnt;
      win *= 4;
    }
    if (g.h_scalars[1] == ~0ull) return 0;
    out[0] = (int64_t)(g.h_scalars[1] >> 8);
    out[1] = out[0] + (int64_t)(g.h_scalars[1] & 0xff);
    return 1;
  }
  ScanParams p{};
  p.text = t->d; p.n = n; p.prog = re->on().d_search; p.filt = make_filter(pl); p.mode = pl.mode;
  p.search_start = start; p.rebase = 0;
  if (pl.mode == MRX_MODE_ANCHORED) {  // dfa.mojo:1888-1892
    if (start != 0) return 0;
    p.mode = MRX_MODE_EVERY_POS;
    launch_attempt_one(p, 0, (int64_t *)(g.scalars + 2), g.stream);
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    int rc = read_scalars(2, 3);
    if (rc) return rc;
    if (!g.h_scalars[4]) return 0;
    out[0] = (int64_t)g.h_scalars[2]; out[1] = (int64_t)g.h_scalars[3];
    return 1;
  }
  if (pl.mode == MRX_MODE_EXACT_LITERAL && start >= n) return 0;  // matcher.mojo:773-774
  const bool at_end = pl.try_at_end && re->prog.start_accepting();
  if (start > n || (start == n && !at_end)) return 0;
  const int64_t origin = start & ~15ll;
  const int64_t last_pos = at_end ? n : n - 1;
  const int64_t tiles = n_tiles_for(origin, last_pos);
  if (tiles <= 0) return 0;
  p.key_lo = start; p.key_hi = n; p.key_hi_inclusive_n = at_end; p.try_at_end = at_end;
  p.out_kind = OUT_FIRST; p.best = g.scalars + 1;
  p.dominated = (re->search_dominated || pl.mode == MRX_MODE_CLASS_RUNS) ? 1 : 0;
  launch_fill_u64(g.scalars + 1, 1, ~0ull, g.stream);
  ++g.launches;
  // windows of growing size so that an early match does not pay for the whole text
  int64_t done = 0, win = 64;
  while (done < tiles) {
    const int64_t cnt = std::min(win, tiles - done);
    p.tile_origin = origin + done * kTile;
    p.n_tiles = (int)cnt;
    launch_scan(p, (int)cnt, g.stream);
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    int rc = read_scalars(1, 1);
    if (rc) return rc;
    if (g.h_scalars[1] != ~0ull) break;
    done += cnt;
    win *= 4;
  }
  if (g.h_scalars[1] == ~0ull) return 0;
  const int64_t pos = (int64_t)g.h_scalars[1];
  launch_attempt_one(p, pos, (int64_t *)(g.scalars + 2), g.stream);
  ++g.launches;
  CUDA_TRY(cudaGetLastError());
  int rc = read_scalars(2, 3);
  if (rc) return rc;
  if (!g.h_scalars[4]) return fail(MRX_E_CUDA, "internal: first-match candidate did not re-verify");
  out[0] = (int64_t)g.h_scalars[2]; out[1] = (int64_t)g.h_scalars[3];
  return 1;
}

int finish_spans(DevSpans &ds, int location, mrx_spans *out) {
  out->n = (size_t)ds.n;
  out->location = location;
  OwnedOut *own = new OwnedOut();
  out->_owner = own;
  if (location == MRX_OUT_DEVICE) {
    own->dev[0] = ds.spans;
    out->start_end = ds.spans;
    return MRX_OK;
  }
  void *h = nullptr;
  int rc = host_alloc(sizeof(int64_t) * 2 * std::max<int64_t>(ds.n, 1), &h);
  if (rc) { dev_release(ds.spans); delete own; out->_owner = nullptr; return rc; }
  own->host[0] = h;
  if (ds.n > 0) {
    CUDA_TRY(cudaMemcpyAsync(h, ds.spans, sizeof(int64_t) * 2 * ds.n, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
  }
  dev_release(ds.spans);
  out->start_end = (int64_t *)h;
  return MRX_OK;
}

// ---- findall over a host buffer, pipelined ---------------------------------------------------------
// The text is uploaded in chunks on the H2D stream; as soon as a chunk has landed, the tiles it completes
// are scanned on the compute stream (one logical job: the ticket, the look-back descriptors and the span
// array continue across launches), and the spans found so far travel back on the D2H stream while the next
// chunk is still arriving.  PCIe is full duplex, so the call costs about one upload of the text.
// Plans served this way read at most kLookAhead bytes past the tile they scan (dense literal sets: none,
// class runs: one byte); every other plan uploads first and scans after.
constexpr size_t kMultiMinBytes = 32u << 20;  // multi-device mode shards a host buffer once every device gets this much
constexpr int64_t kPipeChunk = 32ll << 20;
constexpr int64_t kLookAhead = 16;

cudaEvent_t pipe_event(size_t i) {
  while (g.ev_pool.size() <= i) {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    g.ev_pool.push_back(e);
  }
  return g.ev_pool[i];
}

bool plan_pipelines(const Plan &pl) {
  if (pl.unsupported) return false;
  return pl.dense.valid || pl.mode == MRX_MODE_CLASS_RUNS;
}

int findall_host_pipelined(mrx_regex *re, const uint8_t *host, int64_t n, int64_t core_begin, int64_t core_end,
                           int64_t global_offset, mrx_spans *out) {
  const Plan &pl = re->prog.findall;
  int rc;
  DevTmp d_text, desc;
  DrainSideStreams drain;  // declared after the blocks: destroyed (= streams drained) before they are released
  if ((rc = d_text.get((size_t)n + 512))) return rc;
  uint8_t *dt = d_text.as<uint8_t>();
  mrx_text tview;
  tview.d = dt; tview.n = (size_t)n; tview.owned = false; tview.core_begin = (size_t)core_begin; tview.core_end = (size_t)core_end;
  tview.global_offset = global_offset;
  // chunk = 1/16 of the text, between 4 and 32 MiB: enough stages for the three streams to overlap on
  // small texts, few enough launches and host synchronisations on large ones
  const int64_t chunk = std::min<int64_t>(kPipeChunk, std::max<int64_t>(4ll << 20, ((n / 16) + 0xfffff) & ~0xfffffll));
  const int64_t n_chunks = (n + chunk - 1) / chunk;
  CUDA_TRY(cudaMemsetAsync(dt + n, 0, 512, g.s_in));
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int64_t b = c * chunk, len = std::min(chunk, n - b);
    CUDA_TRY(cudaMemcpyAsync(dt + b, host + b, (size_t)len, cudaMemcpyHostToDevice, g.s_in));
    CUDA_TRY(cudaEventRecord(pipe_event((size_t)c), g.s_in));
  }
  ListJob job{re, &tview, &pl, re->on().d_findall, core_begin, core_end, global_offset};
  if (pl.dense.valid) job.dense = re->on().d_dense_findall;
  if (pl.sand_valid && sand_enabled()) job.sand = re->on().d_sand_findall;
  const int64_t origin = core_begin & ~15ll;
  // dense plans record a match in the tile that holds its end: up to max_len - 1 bytes past the key window
  const int64_t last_pos = pl.dense.valid ? std::min<int64_t>(n - 1, core_end - 1 + std::max(pl.dense.max_len - 1, 0)) : core_end - 1;
  const int64_t tile_bytes = job.sand ? kSandTileBytes : kTile;
  const int64_t tiles = n_tiles_for(origin, last_pos, tile_bytes);
  if (tiles > 0x7fffffff) return fail(MRX_E_INVALID, "text too large for one launch");
  if (tiles <= 0) {
    CUDA_TRY(cudaStreamSynchronize(g.s_in));
    out->n = 0; out->location = MRX_OUT_HOST;
    return MRX_OK;
  }
  const int64_t cap = initial_cap(job);
  void *d_spans = nullptr, *h_spans = nullptr;
  if ((rc = desc.get(sizeof(unsigned long long) * tiles))) return rc;
  if ((rc = dev_alloc(sizeof(int64_t) * 2 * cap, &d_spans))) return rc;
  if ((rc = host_alloc(sizeof(int64_t) * 2 * std::max<int64_t>(cap, 1), &h_spans))) { dev_release(d_spans); return rc; }
  auto cleanup = [&]() { dev_release(d_spans); host_release(h_spans); };
  CUDA_TRY(cudaMemsetAsync(desc.p, 0, sizeof(unsigned long long) * tiles, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.scalars, 0, 8, g.stream));
  const bool runs = !pl.dense.valid;  // CLASS_RUNS: the last run may still be open, its end is written later
  int64_t t0 = 0, copied = 0, total = 0;
  bool overflow = false;
  for (int64_t c = 0; c < n_chunks && !overflow; ++c) {
    const int64_t vis = std::min(n, (c + 1) * chunk);
    const int64_t t1 = (c == n_chunks - 1) ? tiles
                                           : std::min(tiles, std::max<int64_t>(t0, (vis - kLookAhead - origin) / tile_bytes));
    if (t1 <= t0) continue;
    CUDA_TRY(cudaStreamWaitEvent(g.stream, pipe_event((size_t)c), 0));
    CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
    if (g.profiling && c == 0) cudaEventRecord(g.begin_ev(), g.stream);
    if (job.sand) {
      SandParams sp{};
      sp.text = dt; sp.n = n; sp.tile_origin = origin; sp.key_lo = core_begin; sp.key_hi = core_end; sp.table = job.sand;
      sp.n_tiles = (int)t1; sp.tile_base = (int)t0; sp.base0 = (unsigned long long)total; sp.spans = (int64_t *)d_spans;
      sp.cap = cap; sp.rebase = global_offset; sp.desc = desc.as<unsigned long long>(); sp.ticket = g.ticket; sp.total = g.scalars;
      launch_sand(sp, pl.sand.k, pl.sand.tail != 0, (int)(t1 - t0), g.stream);
    } else if (pl.dense.valid) {
      DenseParams dp{};
      dp.text = dt; dp.n = n; dp.tile_origin = origin; dp.key_lo = core_begin; dp.key_hi = core_end; dp.dense = job.dense;
      dp.table_bytes = (int32_t)(pl.dense.next.size() * 2); dp.out_kind = OUT_SPANS;
      dp.n_tiles = (int)t1; dp.tile_base = (int)t0; dp.spans = (int64_t *)d_spans; dp.cap = cap; dp.rebase = global_offset;
      dp.desc = desc.as<unsigned long long>(); dp.ticket = g.ticket; dp.total = g.scalars;
      launch_dense(dp, (int)(t1 - t0), g.stream);
    } else {
      RunsParams p{};
      p.text = dt; p.n = n; p.tile_origin = origin; p.key_lo = core_begin; p.key_hi = core_end; p.prog = re->on().d_findall; p.filt = make_filter(pl);
      p.mode = pl.mode; p.run_min = pl.run_min; p.run_max = pl.run_max; p.n_tiles = (int)t1; p.tile_base = (int)t0;
      p.tile_carry = nullptr; p.spans = (int64_t *)d_spans; p.cap = cap; p.rebase = global_offset;
      p.desc = desc.as<unsigned long long>(); p.ticket = g.ticket; p.total = g.scalars;
      launch_runs(p, (int)(t1 - t0), g.stream);
    }
    if (g.profiling && c == 0) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    if ((rc = read_scalars(0, 1))) { cleanup(); return rc; }  // waits for this launch only; uploads keep going
    total = (int64_t)g.h_scalars[0];
    if (total > cap) { overflow = true; break; }
    const int64_t ready = (t1 == tiles) ? total : (runs ? std::max<int64_t>(total - 1, copied) : total);
    if (ready > copied) {
      CUDA_TRY(cudaMemcpyAsync((int64_t *)h_spans + 2 * copied, (int64_t *)d_spans + 2 * copied,
                               sizeof(int64_t) * 2 * (ready - copied), cudaMemcpyDeviceToHost, g.s_out));
      copied = ready;
    }
    t0 = t1;
  }
  if (overflow) {
    // the capacity guess was too small: finish the upload, then run the plain (resizing) list job
    cleanup();
    CUDA_TRY(cudaStreamSynchronize(g.s_out));
    CUDA_TRY(cudaStreamSynchronize(g.s_in));
    DevSpans ds;
    if ((rc = match_list(re, &tview, pl, re->on().d_findall, ds))) return rc;
    return finish_spans(ds, MRX_OUT_HOST, out);
  }
  note_density(job, total);
  CUDA_TRY(cudaStreamSynchronize(g.s_out));
  CUDA_TRY(cudaStreamSynchronize(g.s_in));
  dev_release(d_spans);
  OwnedOut *own = new OwnedOut();
  own->host[0] = h_spans;
  out->_owner = own;
  out->start_end = (int64_t *)h_spans;
  out->n = (size_t)total;
  out->location = MRX_OUT_HOST;
  return MRX_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int mrx_abi_version(void) { return MRX_ABI_VERSION; }
const char *mrx_last_error(void) { return g_err.c_str(); }

static int init_slot(int device) {  // makes `device` current for this thread and sets its Lib up
  g_cur = &g_devs[device];
  if (g.inited) return MRX_OK;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&g.s_in, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&g.s_out, cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc(&g.ticket, 64));
  CUDA_TRY(cudaMalloc(&g.scalars, 64 * sizeof(unsigned long long)));
  CUDA_TRY(cudaHostAlloc(&g.h_scalars, 64 * sizeof(unsigned long long), cudaHostAllocDefault));
  for (auto &pair : g.evs) { CUDA_TRY(cudaEventCreate(&pair[0])); CUDA_TRY(cudaEventCreate(&pair[1])); }
  g.device = device;
  g.inited = true;
  g.launches = 0;
  return MRX_OK;
}
static void shutdown_slot(int device) {
  Lib *saved = g_cur;
  g_cur = &g_devs[device];
  if (g.inited) {
    cudaSetDevice(device);
    cudaStreamSynchronize(g.stream);
    for (auto &kv : g.dev_free) cudaFree(kv.second);
    for (auto &kv : g.host_free) cudaFreeHost(kv.second);
    g.dev_free.clear();
    g.host_free.clear();
    cudaFree(g.ticket);
    cudaFree(g.scalars);
    cudaFreeHost(g.h_scalars);
    for (auto &pair : g.evs) { cudaEventDestroy(pair[0]); cudaEventDestroy(pair[1]); }
    for (cudaEvent_t e : g.ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(g.s_in);
    cudaStreamDestroy(g.s_out);
    cudaStreamDestroy(g.stream);
    g = Lib();
  }
  g_cur = saved;
}

int mrx_init(int device) {
  if (g_multi <= 1 && g.inited && g.device == device) return MRX_OK;
  mrx_shutdown();
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(MRX_E_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= count || device >= kMaxDevices) return fail(MRX_E_INVALID, "device index out of range");
  g_multi = 0;
  return init_slot(device);
}

int mrx_init_multi(int n_gpus) {
  mrx_shutdown();
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(MRX_E_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (n_gpus < 1 || n_gpus > count || n_gpus > kMaxDevices) return fail(MRX_E_INVALID, "more devices requested than this process can see");
  for (int d = n_gpus - 1; d >= 0; --d) {  // device 0 last: it stays the calling thread's current device
    int rc = init_slot(d);
    if (rc) { mrx_shutdown(); return rc; }
  }
  g_multi = n_gpus;
  return MRX_OK;
}

int mrx_device_count(void) { return g_multi > 1 ? g_multi : (g.inited ? 1 : 0); }

int mrx_shutdown(void) {
  for (int d = 0; d < kMaxDevices; ++d) shutdown_slot(d);
  g_multi = 0;
  g_cur = &g_devs[0];
  return MRX_OK;
}

void *mrx_stream(void) { return g.stream; }
uint64_t mrx_launch_count(void) { return g.launches; }
int mrx_set_profiling(int enabled) {
  g.profiling = enabled != 0;
  g.ev_valid = false;
  g.ev_count = 0;
  return MRX_OK;
}
double mrx_last_kernel_ms(void) {
  if (!g.inited || !g.ev_valid) return -1.0;
  if (cudaEventSynchronize(g.end_ev()) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, g.evs[g.ev_idx % Lib::kProfRing][0], g.end_ev()) != cudaSuccess) return -1.0;
  return (double)ms;
}
// The dominant-kernel times of the last `n` profiled calls, oldest first (one synchronisation, after the loop):
// -> how many were written (at most 63, and no more than were recorded since mrx_set_profiling(1)).
int mrx_kernel_ms_history(double *out_ms, int n) {
  if (!g.inited || !out_ms || n <= 0) return 0;
  if (cudaStreamSynchronize(g.stream) != cudaSuccess) return 0;
  const int have = (int)std::min<unsigned>(g.ev_count, (unsigned)n);
  for (int i = 0; i < have; ++i) {
    const unsigned k = (g.ev_idx - (unsigned)(have - 1 - i)) % Lib::kProfRing;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, g.evs[k][0], g.evs[k][1]) != cudaSuccess) ms = -1.f;
    out_ms[i] = (double)ms;
  }
  return have;
}
int mrx_synchronize(void) {
  int rc = require_device();
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return MRX_OK;
}

// ---- front end ---------------------------------------------------------------------------------
static int regex_from_blob(const std::string &blob, mrx_regex **out) {
  mrx_regex *re = new mrx_regex();
  std::string err;
  if (!load_program((const uint8_t *)blob.data(), blob.size(), re->prog, err)) {
    delete re;
    return fail(MRX_E_INVALID, err);
  }
  re->search_dominated = compute_dominated(re->prog, re->prog.search);
  re->findall_leading_run = compute_leading_run(re->prog, re->prog.findall);
  *out = re;
  return MRX_OK;
}

int mrx_compile(const uint8_t *pattern, size_t n, mrx_regex **out) {
  if (!out || (!pattern && n)) return fail(MRX_E_INVALID, "null argument");
  *out = nullptr;
  RawDfa dfa;
  RegexMeta meta;
  std::string err;
  int rc = compile_frontend(std::string((const char *)pattern, n), dfa, meta, err);
  if (rc != MRX_OK) return fail(rc, err);
  std::string blob;
  if (!pack_blob(dfa, meta, blob, err)) return fail(MRX_E_UNSUPPORTED, err);
  return regex_from_blob(blob, out);
}

int mrx_dfa_pack(const mrx_dfa_desc *desc, const int32_t *transitions, const uint8_t *accepting,
                 const uint8_t *matcher_lookup, const uint8_t *literal, size_t literal_len,
                 const uint8_t *exact_literal, size_t exact_len, const uint8_t *prefilter, size_t prefilter_len,
                 uint8_t *blob, size_t *blob_len) {
  if (!desc || !blob_len || desc->n_states < 0 || (desc->n_states > 0 && (!transitions || !accepting)))
    return fail(MRX_E_INVALID, "null argument");
  RawDfa d;
  d.trans.resize(desc->n_states);
  d.accepting.resize(desc->n_states);
  for (int s = 0; s < desc->n_states; ++s) {
    std::memcpy(d.trans[s].data(), transitions + (size_t)s * 256, 256 * sizeof(int32_t));
    d.accepting[s] = accepting[s] ? 1 : 0;
  }
  d.start_state = desc->start_state;
  d.has_start_anchor = desc->has_start_anchor;
  d.has_end_anchor = desc->has_end_anchor;
  d.is_pure_literal = desc->is_pure_literal;
  d.scan_eligible = desc->scan_eligible;
  if (literal_len) d.literal.assign((const char *)literal, literal_len);
  if (matcher_lookup) { d.has_matcher = true; std::memcpy(d.lookup.data(), matcher_lookup, 256); }
  RegexMeta m;
  m.is_wildcard_any = desc->is_wildcard_any;
  m.is_exact_literal = desc->is_exact_literal;
  m.has_anchors_info = desc->has_anchors_info;
  if (exact_len) m.exact_literal.assign((const char *)exact_literal, exact_len);
  if (prefilter_len) m.prefilter.assign((const char *)prefilter, prefilter_len);
  m.required_byte = desc->required_byte;
  m.fixed_total_width = desc->fixed_sub_total_width;
  m.fixed_num_groups = desc->fixed_sub_num_groups;
  m.fixed_concat = desc->fixed_sub_concat;
  std::memcpy(m.fixed_offsets, desc->fixed_sub_offsets, sizeof m.fixed_offsets);
  std::memcpy(m.fixed_widths, desc->fixed_sub_widths, sizeof m.fixed_widths);
  std::string out, err;
  if (!pack_blob(d, m, out, err)) return fail(MRX_E_UNSUPPORTED, err);
  if (!blob) { *blob_len = out.size(); return MRX_OK; }
  if (*blob_len < out.size()) { *blob_len = out.size(); return fail(MRX_E_INVALID, "blob buffer too small"); }
  std::memcpy(blob, out.data(), out.size());
  *blob_len = out.size();
  return MRX_OK;
}

int mrx_regex_load(const uint8_t *blob, size_t n, mrx_regex **out) {
  if (!blob || !out) return fail(MRX_E_INVALID, "null argument");
  *out = nullptr;
  return regex_from_blob(std::string((const char *)blob, n), out);
}

int mrx_regex_blob(const mrx_regex *re, uint8_t *buf, size_t *len) {
  if (!re || !len) return fail(MRX_E_INVALID, "null argument");
  const std::string &b = re->prog.blob;
  if (!buf) { *len = b.size(); return MRX_OK; }
  if (*len < b.size()) { *len = b.size(); return fail(MRX_E_INVALID, "buffer too small"); }
  std::memcpy(buf, b.data(), b.size());
  *len = b.size();
  return MRX_OK;
}

int mrx_regex_info(const mrx_regex *re, mrx_info *o) {
  if (!re || !o) return fail(MRX_E_INVALID, "null argument");
  const Program &p = re->prog;
  std::memset(o, 0, sizeof *o);
  o->n_states = p.n_states; o->n_classes = p.n_classes; o->start_state = p.start_state;
  o->has_start_anchor = p.flag(BF_START_ANCHOR); o->has_end_anchor = p.flag(BF_END_ANCHOR);
  o->is_pure_literal = p.flag(BF_PURE_LITERAL); o->has_matcher = p.flag(BF_HAS_MATCHER);
  o->scan_eligible = p.flag(BF_SCAN_ELIGIBLE); o->required_byte = p.required_byte;
  o->is_wildcard_any = p.flag(BF_WILDCARD_ANY); o->is_exact_literal = p.flag(BF_EXACT_LITERAL);
  o->has_prefilter = p.flag(BF_HAS_PREFILTER);
  o->findall_mode = p.findall.unsupported ? -1 : p.findall.mode;
  o->search_mode = p.search.unsupported ? -1 : p.search.mode;
  o->findall_direct = p.findall.direct; o->search_direct = p.search.direct;
  o->fixed_sub_total_width = p.fixed_total_width; o->fixed_sub_num_groups = p.fixed_num_groups;
  o->fixed_sub_concat = p.fixed_concat;
  o->findall_dense = p.findall.dense.valid; o->search_dense = p.search.dense.valid;
  o->search_rev_states = p.search.rev.valid ? p.search.rev.n_rev : 0;
  o->search_skip_byte = p.search.rev.valid ? p.search.rev.req_byte : -1;
  o->findall_sand = p.findall.sand_valid ? 1 : 0;
  o->split_fused = p.findall.sep_valid ? 1 : 0;
  o->search_via_prefilter = p.search_via_prefilter ? 1 : 0;
  {
    const Plan &sp = p.search;
    const bool ok = sp.mode == MRX_MODE_UNIFORM_RUN && sp.run_min == sp.run_max && sp.run_min >= 4 && sp.run_min <= 32 &&
                    !p.flag(BF_WILDCARD_ANY);
    o->sub_fused_width = ok ? sp.run_min : 0;
  }
  o->findall_leading_run = re->findall_leading_run ? 1 : 0;
  o->search_dominated = re->search_dominated ? 1 : 0;
  return MRX_OK;
}

void mrx_regex_free(mrx_regex *re) {
  if (!re) return;
  for (RegexOnDevice &o : re->dev) {  // (UVA: a device pointer can be freed from any current device)
    if (o.d_findall) cudaFree(o.d_findall);
    if (o.d_search) cudaFree(o.d_search);
    if (o.d_prefilter) cudaFree(o.d_prefilter);
    if (o.d_dense_findall) cudaFree(o.d_dense_findall);
    if (o.d_dense_search) cudaFree(o.d_dense_search);
    if (o.d_sand_findall) cudaFree(o.d_sand_findall);
    if (o.d_sep_findall) cudaFree(o.d_sep_findall);
    if (o.d_rev_search) cudaFree(o.d_rev_search);
  }
  delete re;
}

// ---- texts ---------------------------------------------------------------------------------------
int mrx_text_upload(const uint8_t *host, size_t n, mrx_text **out) {
  int rc = require_device();
  if (rc) return rc;
  if (!out || (!host && n)) return fail(MRX_E_INVALID, "null argument");
  mrx_text *t = new mrx_text();
  void *d = nullptr;
  rc = dev_alloc(n + 512, &d);  // slack so that 16-byte vector reads past the end stay in bounds
  if (rc) { delete t; return rc; }
  t->d = (uint8_t *)d;
  t->n = n;
  t->owned = true;
  t->core_begin = 0;
  t->core_end = n;
  cudaError_t e = n ? cudaMemcpyAsync(d, host, n, cudaMemcpyHostToDevice, g.stream) : cudaSuccess;
  if (e == cudaSuccess) e = cudaMemsetAsync((uint8_t *)d + n, 0, 512, g.stream);
  if (e != cudaSuccess) {
    cudaStreamSynchronize(g.stream);
    dev_release(d);
    delete t;
    return fail(MRX_E_CUDA, std::string("text upload: ") + cudaGetErrorString(e));
  }
  *out = t;
  return MRX_OK;
}

int mrx_text_wrap_device(const void *dptr, size_t n, mrx_text **out) {
  int rc = require_device();
  if (rc) return rc;
  if (!out || (!dptr && n)) return fail(MRX_E_INVALID, "null argument");
  mrx_text *t = new mrx_text();
  t->d = (uint8_t *)dptr;
  t->n = n;
  t->owned = false;
  t->core_begin = 0;
  t->core_end = n;
  *out = t;
  return MRX_OK;
}

int mrx_text_set_window(mrx_text *t, size_t core_begin, size_t core_end, int64_t global_offset) {
  if (!t || core_begin > core_end || core_end > t->n) return fail(MRX_E_INVALID, "bad window");
  t->core_begin = core_begin;
  t->core_end = core_end;
  t->global_offset = global_offset;
  return MRX_OK;
}

size_t mrx_text_len(const mrx_text *t) { return t ? t->n : 0; }

void mrx_text_free(mrx_text *t) {
  if (!t) return;
  if (t->owned) dev_release(t->d);
  delete t;
}

// ---- matching ------------------------------------------------------------------------------------
int mrx_findall(const mrx_regex *re_c, const mrx_text *t, int location, mrx_spans *out) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !t || !out) return fail(MRX_E_INVALID, "null argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  std::memset(out, 0, sizeof *out);
  if ((rc = ensure_uploaded(re))) return rc;
  DevSpans ds;
  if ((rc = match_list(re, t, re->prog.findall, re->on().d_findall, ds))) return rc;
  return finish_spans(ds, location, out);
}

// findall with the result left in HBM and NO host synchronisation: the kernel is queued and the call returns; the span
// count stays on the device (mrx_spans_count_device: what a stream-ordered consumer or mrx_comm_allgather_device reads)
// until mrx_spans_wait fetches it.  Served this way: plans whose scan kernel writes the final list (class runs,
// chains, literal sets, direct attempt plans); every other plan runs mrx_findall and parks its count on the device.
struct AsyncSpans {
  int64_t cap = 0, keys = 0;
  int density_index = 0;
  mrx_regex *re = nullptr;
  bool pending = false;
  cudaEvent_t done = nullptr;  // recorded after the scan kernel: the blocks go back to the pool only once it has fired
};
static std::map<void *, AsyncSpans> &async_spans() {
  static thread_local std::map<void *, AsyncSpans> m;  // keyed by the result's owner
  return m;
}
int mrx_findall_async(const mrx_regex *re_c, const mrx_text *t, mrx_spans *out) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !t || !out) return fail(MRX_E_INVALID, "null argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  std::memset(out, 0, sizeof *out);
  if ((rc = ensure_uploaded(re))) return rc;
  const Plan &pl = re->prog.findall;
  if (pl.unsupported) return fail(MRX_E_UNSUPPORTED, pl.why);
  const bool single_kernel = pl.mode == MRX_MODE_CLASS_RUNS || pl.mode == MRX_MODE_UNIFORM_RUN ||
                             ((pl.mode == MRX_MODE_LITERAL || pl.mode == MRX_MODE_EXACT_LITERAL || pl.mode == MRX_MODE_FIRST_CLASS ||
                               pl.mode == MRX_MODE_EVERY_POS || pl.mode == MRX_MODE_REQUIRED_BYTE) && pl.direct);
  PoolBlock slot;
  if ((rc = dev_alloc(32, &slot.p))) return rc;
  DevSpans ds;
  AsyncSpans rec;
  rec.re = re;
  rec.keys = (int64_t)t->core_end - (int64_t)t->core_begin;
  if (single_kernel && t->n > 0) {
    ListJob job{re, t, &pl, re->on().d_findall, (int64_t)t->core_begin, (int64_t)t->core_end, t->global_offset};
    if (pl.dense.valid) job.dense = re->on().d_dense_findall;
    if (pl.sand_valid && sand_enabled()) job.sand = re->on().d_sand_findall;
    job.async_total = (unsigned long long *)slot.p;
    job.async_cap = &rec.cap;
    CUDA_TRY(cudaMemsetAsync(slot.p, 0, 32, g.stream));  // (a text with no tiles launches nothing)
    rc = (pl.mode == MRX_MODE_CLASS_RUNS || pl.mode == MRX_MODE_UNIFORM_RUN) ? run_runs_list(job, ds) : run_scan_list(job, false, ds, nullptr);
    if (rc) return rc;
    rec.pending = ds.n < 0;
    if (!rec.pending) ds.n = 0;  // no tiles
  } else {
    if ((rc = match_list(re, t, pl, re->on().d_findall, ds))) return rc;
  }
  if (!rec.pending) {
    g.h_scalars[13] = (unsigned long long)ds.n;
    CUDA_TRY(cudaMemcpyAsync(slot.p, g.h_scalars + 13, 8, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));  // (h_scalars is scratch)
  }
  if ((rc = finish_spans(ds, MRX_OUT_DEVICE, out))) return rc;
  OwnedOut *own = static_cast<OwnedOut *>(out->_owner);
  own->dev[1] = slot.take();
  if (rec.pending) {
    out->n = (size_t)-1;
    OwnerGuard og{own, &out->_owner};  // an error below releases the result instead of leaking it
    CUDA_TRY(cudaEventCreateWithFlags(&rec.done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(rec.done, g.stream));
    og.commit();
    async_spans()[own] = rec;
  }
  return MRX_OK;
}
const void *mrx_spans_count_device(const mrx_spans *s) {
  if (!s || !s->_owner) return nullptr;
  return static_cast<OwnedOut *>(s->_owner)->dev[1];
}
// Fetch the count of an asynchronous findall (one stream synchronisation).  MRX_OK: s->n is the number of spans.
// 1: the span array was too small for the text (the density estimate of the previous call was off): the caller
// repeats the call -- the estimate has been corrected.
int mrx_spans_wait(mrx_spans *s) {
  if (!s || !s->_owner) return fail(MRX_E_INVALID, "null argument");
  if (s->n != (size_t)-1) return MRX_OK;
  OwnedOut *own = static_cast<OwnedOut *>(s->_owner);
  auto it = async_spans().find(own);
  if (it == async_spans().end()) return fail(MRX_E_INVALID, "not an asynchronous result of this thread");
  const AsyncSpans rec = it->second;
  async_spans().erase(it);
  if (rec.done) cudaEventDestroy(rec.done);
  CUDA_TRY(cudaMemcpyAsync(g.h_scalars + 13, own->dev[1], 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  const int64_t total = (int64_t)g.h_scalars[13];
  if (rec.keys > 0) rec.re->on().density[0] = (double)total / (double)rec.keys;
  if (total > rec.cap) { s->n = (size_t)rec.cap; return 1; }
  s->n = (size_t)total;
  return MRX_OK;
}

static int findall_host_shard_single(const mrx_regex *re_c, const uint8_t *host, size_t n, size_t core_begin, size_t core_end,
                                     int64_t global_offset, mrx_spans *out) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !out || (!host && n) || core_begin > core_end || core_end > n) return fail(MRX_E_INVALID, "bad argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  std::memset(out, 0, sizeof *out);
  if ((rc = ensure_uploaded(re))) return rc;
  if (plan_pipelines(re->prog.findall) && (int64_t)n >= (16ll << 20))
    return findall_host_pipelined(re, host, (int64_t)n, (int64_t)core_begin, (int64_t)core_end, global_offset, out);
  mrx_text *t = nullptr;
  if ((rc = mrx_text_upload(host, n, &t))) return rc;
  if ((rc = mrx_text_set_window(t, core_begin, core_end, global_offset)) == MRX_OK)
    rc = mrx_findall(re_c, t, MRX_OUT_HOST, out);
  mrx_text_free(t);
  return rc;
}

int mrx_findall_host_shard(const mrx_regex *re_c, const uint8_t *host, size_t n, size_t core_begin, size_t core_end,
                           int64_t global_offset, mrx_spans *out) {
  return findall_host_shard_single(re_c, host, n, core_begin, core_end, global_offset, out);
}

static int findall_host_multi(const mrx_regex *re_c, const uint8_t *host, size_t n, mrx_spans *out);
static bool findall_shardable(const mrx_regex *re);

int mrx_findall_host(const mrx_regex *re_c, const uint8_t *host, size_t n, mrx_spans *out) {
  if (g_multi > 1 && re_c && n >= (size_t)g_multi * kMultiMinBytes && findall_shardable(re_c))
    return findall_host_multi(re_c, host, n, out);
  return findall_host_shard_single(re_c, host, n, 0, n, 0, out);
}

int mrx_search(const mrx_regex *re_c, const mrx_text *t, int64_t start, int64_t out[2]) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !t || !out) return fail(MRX_E_INVALID, "null argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  if ((rc = ensure_uploaded(re))) return rc;
  return search_impl(re, t, start, out);
}

int mrx_match_first(const mrx_regex *re_c, const mrx_text *t, int64_t start, int64_t out[2]) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !t || !out) return fail(MRX_E_INVALID, "null argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  const Program &p = re->prog;
  const int64_t n = (int64_t)t->n;
  if (p.flag(BF_WILDCARD_ANY)) {  // matcher.mojo:741-745
    if (start <= n) { out[0] = start; out[1] = n; return 1; }
    return 0;
  }
  if ((rc = ensure_uploaded(re))) return rc;
  if (p.flag(BF_START_ANCHOR) && start > 0) return 0;  // dfa.mojo:1866-1867
  if (start > n || start < 0) return 0;
  // _try_match_at_position(require_exact_position=True): literal verify or one attempt
  Plan one = p.findall;  // only the attempt semantics matter here
  ScanParams sp{};
  sp.text = t->d; sp.n = n; sp.rebase = 0;
  DevTmp tmpprog;
  if (p.flag(BF_PURE_LITERAL)) {
    sp.mode = MRX_MODE_LITERAL;
    sp.prog = (p.findall.mode == MRX_MODE_LITERAL) ? re->on().d_findall : re->on().d_search;
    if (p.findall.mode != MRX_MODE_LITERAL && p.search.mode != MRX_MODE_LITERAL)
      return fail(MRX_E_UNSUPPORTED, "match_first on a literal routed through the exact-literal path");
  } else {
    sp.mode = MRX_MODE_EVERY_POS;
    // any uploaded program of this regex carries the same automaton and shortcut flag semantics,
    // except that the shortcut depends only on the engine -> take the findall program
    sp.prog = re->on().d_findall;
    (void)one;
  }
  launch_attempt_one(sp, start, (int64_t *)(g.scalars + 2), g.stream);
  ++g.launches;
  CUDA_TRY(cudaGetLastError());
  if ((rc = read_scalars(2, 3))) return rc;
  if (!g.h_scalars[4]) return 0;
  out[0] = (int64_t)g.h_scalars[2];
  out[1] = (int64_t)g.h_scalars[3];
  return 1;
}

int mrx_search_batch(const mrx_regex *re_c, const mrx_text *t, const int64_t *doc_offsets, size_t n_docs,
                     int64_t doc_stride, int location, int64_t *out_start_end) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !t || !out_start_end || (!doc_offsets && doc_stride <= 0)) return fail(MRX_E_INVALID, "bad argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  if ((rc = ensure_uploaded(re))) return rc;
  const Plan &pl = re->prog.search;
  if (pl.unsupported) return fail(MRX_E_UNSUPPORTED, pl.why);
  if (n_docs == 0) return MRX_OK;
  const int64_t n = (int64_t)t->n;
  if (!doc_offsets && (int64_t)(n_docs - 1) > n / doc_stride) return fail(MRX_E_INVALID, "more fixed-size documents than the text holds");
  const bool trivial = pl.mode == MRX_MODE_WILDCARD_ANY || pl.mode == MRX_MODE_ANCHORED;
  // offsets / results on the device
  DevTmp d_offs, d_best, d_out;
  const int64_t *offs_dev = nullptr;
  if (doc_offsets) {
    if (location == MRX_OUT_HOST) {
      // host offsets are validated here (a bad table would send the kernels out of bounds)
      if (doc_offsets[0] < 0 || doc_offsets[n_docs] > n) return fail(MRX_E_INVALID, "doc_offsets outside the text");
      for (size_t i = 0; i < n_docs; ++i)
        if (doc_offsets[i] > doc_offsets[i + 1]) return fail(MRX_E_INVALID, "doc_offsets must be non-decreasing");
      if ((rc = d_offs.get(sizeof(int64_t) * (n_docs + 1)))) return rc;
      CUDA_TRY(cudaMemcpyAsync(d_offs.p, doc_offsets, sizeof(int64_t) * (n_docs + 1), cudaMemcpyHostToDevice, g.stream));
      offs_dev = d_offs.as<int64_t>();
    } else {
      offs_dev = doc_offsets;
    }
  }
  if (trivial) {  // `.*`: every document whole; `^...`: one attempt at the document start
    int64_t *outp = out_start_end;
    if (location == MRX_OUT_HOST) {
      if ((rc = d_out.get(sizeof(int64_t) * 2 * n_docs))) return rc;
      outp = d_out.as<int64_t>();
    }
    ScanParams p{};
    p.text = t->d; p.n = n; p.prog = re->on().d_search; p.mode = MRX_MODE_EVERY_POS;
    p.doc_offsets = offs_dev; p.doc_stride = doc_stride; p.n_docs = (int64_t)n_docs;
    launch_batch_trivial(p, pl.mode == MRX_MODE_WILDCARD_ANY ? 1 : 0, outp, g.stream);
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    if (location == MRX_OUT_HOST) {
      CUDA_TRY(cudaMemcpyAsync(out_start_end, outp, sizeof(int64_t) * 2 * n_docs, cudaMemcpyDeviceToHost, g.stream));
      CUDA_TRY(cudaStreamSynchronize(g.stream));
    }
    return MRX_OK;
  }
  if (re->on().d_rev_search && re->on().d_search && !getenv("MRX_NO_LINES") && !re->prog.search_via_prefilter) {
    // reverse-automaton kernel: one thread per document, results written in place
    int64_t *outp = out_start_end;
    if (location == MRX_OUT_HOST) {
      if ((rc = d_out.get(sizeof(int64_t) * 2 * n_docs))) return rc;
      outp = d_out.as<int64_t>();
    }
    LinesParams lp{};
    lp.text = t->d; lp.n = n; lp.tables = re->on().d_rev_search;
    lp.rev_bytes = (int32_t)(pl.rev.rev.size() * 2); lp.fwd_bytes = (int32_t)(pl.rev.fwd.size() * 2);
    lp.fwd_start = pl.rev.fwd_start; lp.doc_offsets = offs_dev; lp.doc_stride = doc_stride;
    lp.n_docs = (int64_t)n_docs; lp.out = outp; lp.req_byte = pl.rev.req_byte; lp.ticket = g.ticket;
    CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
    if (g.profiling && !g.prof_outer) cudaEventRecord(g.begin_ev(), g.stream);
    launch_lines(lp, g.stream);
    if (g.profiling && !g.prof_outer) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    if (location == MRX_OUT_HOST) {
      CUDA_TRY(cudaMemcpyAsync(out_start_end, outp, sizeof(int64_t) * 2 * n_docs, cudaMemcpyDeviceToHost, g.stream));
      CUDA_TRY(cudaStreamSynchronize(g.stream));
    }
    // MRX_OUT_DEVICE: the results are ordered on the library stream (mrx_stream / mrx_synchronize)
    return MRX_OK;
  }
  if ((rc = d_best.get(sizeof(unsigned long long) * n_docs))) return rc;
  launch_fill_u64(d_best.as<unsigned long long>(), (int64_t)n_docs, ~0ull, g.stream);
  ++g.launches;
  const int64_t tiles = n_tiles_for(0, n - 1);
  DevTmp d_lit;
  if (re->prog.search_via_prefilter) {
    // matcher.mojo:784-796 per document: the first occurrence of the prefilter literal, then the search from there
    const Plan &pre = re->prog.prefilter_plan;
    if ((rc = d_lit.get(sizeof(unsigned long long) * n_docs))) return rc;
    launch_fill_u64(d_lit.as<unsigned long long>(), (int64_t)n_docs, ~0ull, g.stream);
    ++g.launches;
    if (tiles > 0) {
      ScanParams q{};
      q.text = t->d; q.n = n; q.prog = re->on().d_prefilter; q.filt = make_filter(pre); q.mode = pre.mode;
      q.tile_origin = 0; q.key_lo = 0; q.key_hi = n; q.search_start = -1; q.n_tiles = (int)tiles;
      q.out_kind = OUT_BATCH; q.doc_offsets = offs_dev; q.doc_stride = doc_stride; q.n_docs = (int64_t)n_docs;
      q.doc_best = d_lit.as<unsigned long long>();
      launch_scan(q, (int)tiles, g.stream);
      ++g.launches;
    }
  }
  ScanParams p{};
  p.text = t->d; p.n = n; p.prog = re->on().d_search; p.filt = make_filter(pl); p.mode = pl.mode;
  p.tile_origin = 0; p.key_lo = 0; p.key_hi = n; p.search_start = -1;
  p.out_kind = OUT_BATCH; p.doc_offsets = offs_dev; p.doc_stride = doc_stride; p.n_docs = (int64_t)n_docs;
  p.doc_best = d_best.as<unsigned long long>();
  p.doc_min = re->prog.search_via_prefilter ? d_lit.as<unsigned long long>() : nullptr;
  p.dominated = (re->search_dominated || pl.mode == MRX_MODE_CLASS_RUNS) ? 1 : 0;
  if (re->prog.search_via_prefilter) p.dominated = 0;  // a dominated candidate may be the first one at or after the literal
  if (tiles > 0) {
    p.n_tiles = (int)tiles;
    if (g.profiling && !g.prof_outer) cudaEventRecord(g.begin_ev(), g.stream);
    launch_scan(p, (int)tiles, g.stream);
    if (g.profiling && !g.prof_outer) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
    ++g.launches;
  }
  int64_t *outp = out_start_end;
  if (location == MRX_OUT_HOST) {
    if ((rc = d_out.get(sizeof(int64_t) * 2 * n_docs))) return rc;
    outp = d_out.as<int64_t>();
  }
  launch_batch_finalize(d_best.as<unsigned long long>(), (int64_t)n_docs, offs_dev, doc_stride, n,
                        (pl.try_at_end && re->prog.start_accepting()) ? 1 : 0, outp, g.stream);
  ++g.launches;
  CUDA_TRY(cudaGetLastError());
  if (location == MRX_OUT_HOST) {
    CUDA_TRY(cudaMemcpyAsync(out_start_end, outp, sizeof(int64_t) * 2 * n_docs, cudaMemcpyDeviceToHost, g.stream));
  }
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return MRX_OK;
}

// replacement template: matcher.mojo:1437-1482 (_parse_repl_template, _has_group_refs)
static bool repl_has_group_refs(const uint8_t *r, size_t n) {
  for (size_t i = 0; i + 1 < n; ++i)
    if (r[i] == '\\' && r[i + 1] >= '1' && r[i + 1] <= '9') return true;
  return false;
}

// sub() in one pass (k_subst): every match of the plan is a run of exactly W bytes of one class, every replacement
// R bytes long
static int subst_stage_bytes(int W, int R) {
  const int grow = R > W ? (R - W) * (kTile / W + 1) : 0;
  return (kTile + grow + 96 + 127) & ~127;  // + the bytes of a match that reaches past the tile, + the copy-out's read-ahead
}
static bool sub_fused_applies(const Program &p, const std::vector<int32_t> &flat) {
  const Plan &pl = p.search;
  if (pl.mode != MRX_MODE_UNIFORM_RUN || pl.run_min != pl.run_max) return false;
  const int W = pl.run_min, R = (int)flat.size();
  if (W < 4 || W > 32 || R > kSubstMaxRepl || p.flag(BF_WILDCARD_ANY)) return false;
  for (int32_t src : flat)
    if (src >= W) return false;
  return subst_stage_bytes(W, R) <= kSubstMaxStage;
}
static int sub_fused(mrx_regex *re, const mrx_text *t, const std::vector<int32_t> &flat, int location, mrx_bytes *out,
                     OwnedOut *own) {
  const Plan &pl = re->prog.search;
  const int64_t n = (int64_t)t->n;
  const int W = pl.run_min, R = (int)flat.size();
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffff) return fail(MRX_E_INVALID, "text too large for one launch");
  int rc;
  DevTmp desc, trail, allf, carry, d_tpl;
  if ((rc = desc.get(sizeof(unsigned long long) * tiles))) return rc;
  if ((rc = trail.get(sizeof(long long) * tiles))) return rc;
  if ((rc = allf.get(sizeof(int) * tiles))) return rc;
  if ((rc = carry.get(sizeof(long long) * tiles))) return rc;
  if ((rc = d_tpl.get(sizeof(int32_t) * (kSubstMaxRepl + 4)))) return rc;
  if (R) CUDA_TRY(cudaMemcpyAsync(d_tpl.p, flat.data(), sizeof(int32_t) * R, cudaMemcpyHostToDevice, g.stream));
  // every match grows the text by R - W bytes and there is at most one match per W bytes
  const int64_t out_bound = n + (R > W ? (int64_t)(R - W) * (n / W) : 0);
  void *d_out = nullptr;
  if ((rc = dev_alloc((size_t)out_bound + 32, &d_out))) return rc;
  own->dev[0] = d_out;
  CUDA_TRY(cudaMemsetAsync(desc.p, 0, sizeof(unsigned long long) * tiles, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.scalars, 0, 16, g.stream));
  ProfOuter prof;
  prof.arm();
  launch_tile_trailing(t->d, n, 0, (int)tiles, re->on().d_search, trail.as<long long>(), allf.as<int>(), g.stream);
  launch_carry_scan((int)tiles, trail.as<long long>(), allf.as<int>(), carry.as<long long>(), nullptr, g.stream);
  SubstParams sp{};
  sp.text = t->d; sp.n = n; sp.prog = re->on().d_search; sp.filt = make_filter(pl); sp.width = W; sp.repl_len = R;
  sp.tpl = d_tpl.as<int32_t>(); sp.n_tiles = (int)tiles; sp.stage_bytes = subst_stage_bytes(W, R);
  sp.tile_carry = carry.as<long long>(); sp.out = (uint8_t *)d_out; sp.desc = desc.as<unsigned long long>();
  sp.ticket = g.ticket; sp.totals = g.scalars;
  launch_subst(sp, g.stream);
  if (g.profiling) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
  prof.disarm();
  g.launches += 3;
  CUDA_TRY(cudaGetLastError());
  if ((rc = read_scalars(0, 2))) return rc;
  const int64_t total = (int64_t)g.h_scalars[0];
  if (total > out_bound) return fail(MRX_E_CUDA, "fused sub wrote past its bound");
  out->n = (size_t)total;
  if (location == MRX_OUT_DEVICE) {
    out->bytes = (uint8_t *)d_out;
    return MRX_OK;
  }
  void *h = nullptr;
  if ((rc = host_alloc((size_t)total + 1, &h))) return rc;
  own->host[0] = h;
  if (total) CUDA_TRY(cudaMemcpyAsync(h, d_out, (size_t)total, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  dev_release(d_out);
  own->dev[0] = nullptr;
  out->bytes = (uint8_t *)h;
  return MRX_OK;
}

int mrx_sub(const mrx_regex *re_c, const uint8_t *repl, size_t repl_n, int64_t count, const mrx_text *t,
            int location, mrx_bytes *out) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !t || !out || (!repl && repl_n)) return fail(MRX_E_INVALID, "null argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  std::memset(out, 0, sizeof *out);
  out->location = location;
  const Program &p = re->prog;
  const int64_t n = (int64_t)t->n;
  if (t->core_begin != 0 || t->core_end != t->n) return fail(MRX_E_INVALID, "sub on a windowed text: assemble per shard on whole buffers");
  // template
  SubTemplate tpl{};
  int64_t repl_len = 0;
  const bool use_groups = repl_has_group_refs(repl, repl_n);
  if (use_groups) {
    if (p.fixed_total_width < 0)
      return fail(MRX_E_UNSUPPORTED, "capture-group references need the reference's NFA group engine (matcher.mojo:1783-1821)");
    size_t i = 0, lit_start = 0;
    auto push_lit = [&](size_t a, size_t b) -> bool {
      if (b <= a) return true;
      if (tpl.n_segs >= kMaxTplSegs) return false;
      tpl.group_off[tpl.n_segs] = -1; tpl.lit_off[tpl.n_segs] = (int32_t)a; tpl.len[tpl.n_segs] = (int32_t)(b - a);
      ++tpl.n_segs;
      return true;
    };
    while (i < repl_n) {
      if (repl[i] == '\\' && i + 1 < repl_n && repl[i + 1] >= '1' && repl[i + 1] <= '9') {
        if (!push_lit(lit_start, i)) return fail(MRX_E_UNSUPPORTED, "replacement template has too many segments");
        const int gref = repl[i + 1] - '0';
        if (gref <= p.fixed_num_groups) {  // refs past the last group expand to nothing (matcher.mojo:1613-1620)
          if (tpl.n_segs >= kMaxTplSegs) return fail(MRX_E_UNSUPPORTED, "replacement template has too many segments");
          tpl.group_off[tpl.n_segs] = p.fixed_offsets[gref]; tpl.lit_off[tpl.n_segs] = 0; tpl.len[tpl.n_segs] = p.fixed_widths[gref];
          ++tpl.n_segs;
        }
        i += 2;
        lit_start = i;
        continue;
      }
      ++i;
    }
    if (!push_lit(lit_start, repl_n)) return fail(MRX_E_UNSUPPORTED, "replacement template has too many segments");
  } else if (repl_n) {
    tpl.n_segs = 1; tpl.group_off[0] = -1; tpl.lit_off[0] = 0; tpl.len[0] = (int32_t)repl_n;
  }
  for (int k = 0; k < tpl.n_segs; ++k) repl_len += tpl.len[k];
  OwnedOut *own = new OwnedOut();
  out->_owner = own;
  OwnerGuard og{own, &out->_owner};
  auto bail = [&](int code) { return code; };  // the guard releases the owner
  if (n == 0) {  // matcher.mojo:1697-1698
    out->n = 0;
    out->bytes = nullptr;
    og.commit();
    return MRX_OK;
  }
  if ((rc = ensure_uploaded(re))) return bail(rc);
  // flattened replacement: one int32 per output byte of the template (see k_assemble, k_subst)
  std::vector<int32_t> flat;
  flat.reserve((size_t)repl_len);
  for (int k = 0; k < tpl.n_segs; ++k)
    for (int i = 0; i < tpl.len[k]; ++i)
      flat.push_back(tpl.group_off[k] >= 0 ? tpl.group_off[k] + i : -1 - (int32_t)repl[tpl.lit_off[k] + i]);
  if (count <= 0 && sub_fused_applies(p, flat) && !getenv("MRX_NO_FUSED_SUB")) {
    // a fixed-width chain over one class (config 4): finder, template expansion and output in one kernel
    if ((rc = sub_fused(re, t, flat, location, out, own))) return bail(rc);
    og.commit();
    return MRX_OK;
  }
  DevSpans ms;
  ProfOuter prof;
  prof.arm();
  // the assembly works in buffer coordinates: a shard's global offset (mrx_text_set_window) applies to the
  // spans findall reports, never to the match list sub / split are built from
  mrx_text local = *t;
  local.global_offset = 0;
  t = &local;
  if (p.flag(BF_WILDCARD_ANY)) {
    // `.*`: the replace loop (matcher.mojo:1824-1847) calls match_next(text, pos) until it fails: (0, n), then
    // the empty match (n, n) at the end of the text (matcher.mojo:762-766) -- two replacements, not the one
    // span match_all reports
    PoolBlock sp;
    if ((rc = dev_alloc(32, &sp.p))) return bail(rc);
    const int64_t h[4] = {0, n, n, n};
    CUDA_TRY(cudaMemcpyAsync(sp.p, h, 32, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    ms.spans = (int64_t *)sp.take();
    ms.n = 2;
    rc = MRX_OK;
  } else
  rc = match_list(re, t, p.search, re->on().d_search, ms);
  prof.disarm();
  if (rc) return bail(rc);
  int64_t M = ms.n;
  if (count > 0 && M > count) M = count;
  DevTmp d_tpl, d_offs, d_desc, d_bad, d_tseg;
  struct SpansGuard { int64_t *p; ~SpansGuard() { dev_release(p); } } guard{ms.spans};
  if ((rc = d_tpl.get(sizeof(int32_t) * (flat.size() + 4)))) return bail(rc);
  if (!flat.empty()) CUDA_TRY(cudaMemcpyAsync(d_tpl.p, flat.data(), sizeof(int32_t) * flat.size(), cudaMemcpyHostToDevice, g.stream));
  const int64_t seg_tiles = (M + 1 + kSegTile - 1) / kSegTile;
  const int64_t out_bound = n + M * (repl_len + 1) + 1;
  if ((rc = d_offs.get(sizeof(int64_t) * (M + 2)))) return bail(rc);
  if ((rc = d_desc.get(sizeof(unsigned long long) * seg_tiles))) return bail(rc);
  if ((rc = d_bad.get(16))) return bail(rc);
  if ((rc = d_tseg.get(sizeof(int64_t) * (out_bound / kAsmTile + 2)))) return bail(rc);
  CUDA_TRY(cudaMemsetAsync(d_desc.p, 0, sizeof(unsigned long long) * seg_tiles, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
  CUDA_TRY(cudaMemsetAsync(d_bad.p, 0, 16, g.stream));
  launch_seg_lengths(ms.spans, M, n, SEG_SUB, repl_len, d_desc.as<unsigned long long>(), g.ticket, d_offs.as<int64_t>(),
                     d_bad.as<int>(), d_tseg.as<int64_t>(), g.stream);
  ++g.launches;
  CUDA_TRY(cudaGetLastError());
  int64_t total = 0;
  CUDA_TRY(cudaMemcpyAsync(g.h_scalars + 8, d_offs.as<int64_t>() + (M + 1), 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  total = (int64_t)g.h_scalars[8];
  void *d_out = nullptr;
  if ((rc = dev_alloc((size_t)total + 16, &d_out))) return bail(rc);
  own->dev[0] = d_out;
  {
    AsmParams ap{};
    ap.text = t->d; ap.n = n; ap.spans = ms.spans; ap.M = M; ap.offs = d_offs.as<int64_t>();
    ap.tile_seg = d_tseg.as<int64_t>(); ap.n_tiles = (total + kAsmTile - 1) / kAsmTile; ap.mode = SEG_SUB;
    ap.repl_len = repl_len; ap.tpl_src = d_tpl.as<int32_t>(); ap.total = total; ap.out = (uint8_t *)d_out;
    launch_assemble(ap, g.stream);
    g.launches += total > 0 ? 1 : 0;
    if (g.profiling) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
  }
  CUDA_TRY(cudaGetLastError());
  out->n = (size_t)total;
  if (location == MRX_OUT_DEVICE) {
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    out->bytes = (uint8_t *)d_out;
    og.commit();
    return MRX_OK;
  }
  void *h = nullptr;
  if ((rc = host_alloc((size_t)total + 1, &h))) return bail(rc);
  own->host[0] = h;
  if (total) CUDA_TRY(cudaMemcpyAsync(h, d_out, (size_t)total, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  dev_release(d_out);
  own->dev[0] = nullptr;
  out->bytes = (uint8_t *)h;
  og.commit();
  return MRX_OK;
}

// split() for X* c Y* patterns in one pass (k_split): no match list.  *done = false when the text holds a run the
// kernel's windows cannot measure (the caller then takes the general path).
static int split_fused(mrx_regex *re, const mrx_text *t, int location, mrx_pieces *out, bool *done) {
  *done = false;
  const int64_t n = (int64_t)t->n;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffff) return fail(MRX_E_INVALID, "text too large for one launch");
  int rc;
  DevTmp desc;
  if ((rc = desc.get(sizeof(unsigned long long) * tiles))) return rc;
  OwnedOut *own = new OwnedOut();
  out->_owner = own;
  OwnerGuard og{own, &out->_owner};
  void *d_bytes = nullptr;
  if ((rc = dev_alloc((size_t)n + 32, &d_bytes))) return rc;
  own->dev[0] = d_bytes;
  const double dens = re->on().density[0];
  int64_t cap = dens > 0.0 ? (int64_t)(dens * 1.125 * (double)n) + 4096 : std::min<int64_t>(n / 6 + 1024, 1ll << 27);
  ProfOuter prof;
  for (int attempt = 0; attempt < 3; ++attempt) {
    void *d_offs = nullptr;
    if ((rc = dev_alloc(sizeof(int64_t) * (size_t)(cap + 2), &d_offs))) return rc;
    own->dev[1] = d_offs;
    CUDA_TRY(cudaMemsetAsync(desc.p, 0, sizeof(unsigned long long) * tiles, g.stream));
    CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
    CUDA_TRY(cudaMemsetAsync(g.scalars, 0, 32, g.stream));
    CUDA_TRY(cudaMemsetAsync(d_offs, 0, 8, g.stream));
    SplitParams sp{};
    sp.text = t->d; sp.n = n; sp.table = re->on().d_sep_findall; sp.n_tiles = (int)tiles; sp.out = (uint8_t *)d_bytes;
    sp.offs = (int64_t *)d_offs; sp.offs_cap = cap; sp.desc = desc.as<unsigned long long>(); sp.ticket = g.ticket;
    sp.totals = g.scalars; sp.flags = reinterpret_cast<int *>(g.scalars + 2);
    prof.arm();
    launch_split(sp, g.stream);
    if (g.profiling) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
    prof.disarm();
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    if ((rc = read_scalars(0, 3))) return rc;
    const int64_t total = (int64_t)g.h_scalars[0], M = (int64_t)g.h_scalars[1];
    const int flags = (int)(g.h_scalars[2] & 0xffffffffu);
    if (flags & 1) return MRX_OK;  // a long run: general path (*done stays false; the guard releases the buffers)
    if (flags & 2)
      return fail(MRX_E_UNDEFINED, "separator spans overlap (text[prev_end:start] with prev_end > start): the reference's "
                                   "behaviour is undefined here (matcher.mojo:1389)");
    re->on().density[0] = (double)M / (double)n;
    if (M > cap) {  // the offsets array was too small: once more with the exact size
      dev_release(d_offs);
      own->dev[1] = nullptr;
      cap = M;
      continue;
    }
    // offs[M + 1] = total (offs[0] = 0 is already there)
    g.h_scalars[12] = (unsigned long long)total;
    CUDA_TRY(cudaMemcpyAsync((int64_t *)d_offs + (M + 1), g.h_scalars + 12, 8, cudaMemcpyHostToDevice, g.stream));
    out->n = (size_t)(M + 1);
    out->location = location;
    *done = true;
    if (location == MRX_OUT_DEVICE) {
      CUDA_TRY(cudaStreamSynchronize(g.stream));
      out->bytes = (uint8_t *)d_bytes;
      out->offsets = (int64_t *)d_offs;
      og.commit();
      return MRX_OK;
    }
    void *hb = nullptr, *ho = nullptr;
    if ((rc = host_alloc((size_t)total + 1, &hb))) return rc;
    own->host[0] = hb;
    if ((rc = host_alloc(sizeof(int64_t) * (M + 2), &ho))) return rc;
    own->host[1] = ho;
    if (total) CUDA_TRY(cudaMemcpyAsync(hb, d_bytes, (size_t)total, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaMemcpyAsync(ho, d_offs, sizeof(int64_t) * (M + 2), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    dev_release(d_bytes); own->dev[0] = nullptr;
    dev_release(d_offs); own->dev[1] = nullptr;
    out->bytes = (uint8_t *)hb;
    out->offsets = (int64_t *)ho;
    og.commit();
    return MRX_OK;
  }
  return fail(MRX_E_CUDA, "separator count did not stabilise");
}

int mrx_split(const mrx_regex *re_c, const mrx_text *t, int64_t maxsplit, int location, mrx_pieces *out) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !t || !out) return fail(MRX_E_INVALID, "null argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  std::memset(out, 0, sizeof *out);
  out->location = location;
  const int64_t n = (int64_t)t->n;
  if (t->core_begin != 0 || t->core_end != t->n) return fail(MRX_E_INVALID, "split on a windowed text: assemble per shard on whole buffers");
  if ((rc = ensure_uploaded(re))) return rc;
  if (re->prog.findall.sep_valid && maxsplit == 0 && n > 0 && re->on().d_sep_findall && !getenv("MRX_NO_FUSED_SPLIT")) {
    // X* c Y* (config 5): finder, compaction and piece offsets in one kernel; texts with runs the kernel's windows
    // cannot measure fall through to the general path
    bool done = false;
    rc = split_fused(re, t, location, out, &done);
    if (rc || done) return rc;
    std::memset(out, 0, sizeof *out);
    out->location = location;
  }
  DevSpans ms;
  ProfOuter prof;
  prof.arm();
  mrx_text local = *t;  // buffer coordinates for the assembly (see mrx_sub)
  local.global_offset = 0;
  t = &local;
  rc = match_list(re, t, re->prog.findall, re->on().d_findall, ms);
  prof.disarm();
  if (rc) return rc;
  struct SpansGuard { int64_t *p; ~SpansGuard() { dev_release(p); } } guard{ms.spans};
  int64_t M = ms.n;
  if (maxsplit < 0) M = 0;  // matcher.mojo:1386-1387 with splits_done(0) >= maxsplit
  else if (maxsplit > 0 && M > maxsplit) M = maxsplit;
  OwnedOut *own = new OwnedOut();
  out->_owner = own;
  OwnerGuard og{own, &out->_owner};
  auto bail = [&](int code) { return code; };  // the guard releases the owner
  DevTmp d_desc, d_bad, d_tseg;
  const int64_t seg_tiles = (M + 1 + kSegTile - 1) / kSegTile;
  void *d_offs = nullptr;
  if ((rc = dev_alloc(sizeof(int64_t) * (M + 2), &d_offs))) return bail(rc);
  own->dev[1] = d_offs;
  if ((rc = d_desc.get(sizeof(unsigned long long) * seg_tiles))) return bail(rc);
  if ((rc = d_bad.get(16))) return bail(rc);
  if ((rc = d_tseg.get(sizeof(int64_t) * (n / kAsmTile + 2)))) return bail(rc);
  CUDA_TRY(cudaMemsetAsync(d_desc.p, 0, sizeof(unsigned long long) * seg_tiles, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.ticket, 0, 16, g.stream));
  CUDA_TRY(cudaMemsetAsync(d_bad.p, 0, 16, g.stream));
  if ((rc = debug_sync("split: match list"))) return bail(rc);
  launch_seg_lengths(ms.spans, M, n, SEG_SPLIT, 0, d_desc.as<unsigned long long>(), g.ticket, (int64_t *)d_offs,
                     d_bad.as<int>(), d_tseg.as<int64_t>(), g.stream);
  ++g.launches;
  if ((rc = debug_sync("split: k_seg_lengths"))) return bail(rc);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(g.h_scalars + 8, (int64_t *)d_offs + (M + 1), 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaMemcpyAsync(g.h_scalars + 9, d_bad.p, 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  if ((int)(g.h_scalars[9] & 0xffffffffu) != 0)
    return bail(fail(MRX_E_UNDEFINED, "separator spans overlap (text[prev_end:start] with prev_end > start): the reference's "
                                      "behaviour is undefined here (matcher.mojo:1389)"));
  const int64_t total = (int64_t)g.h_scalars[8];
  void *d_bytes = nullptr;
  if ((rc = dev_alloc((size_t)total + 16, &d_bytes))) return bail(rc);
  own->dev[0] = d_bytes;
  {
    AsmParams ap{};
    ap.text = t->d; ap.n = n; ap.spans = ms.spans; ap.M = M; ap.offs = (int64_t *)d_offs;
    ap.tile_seg = d_tseg.as<int64_t>(); ap.n_tiles = (total + kAsmTile - 1) / kAsmTile; ap.mode = SEG_SPLIT;
    ap.repl_len = 0; ap.tpl_src = nullptr; ap.total = total; ap.out = (uint8_t *)d_bytes;
    launch_assemble(ap, g.stream);
    if ((rc = debug_sync("split: k_assemble"))) return bail(rc);
    g.launches += total > 0 ? 1 : 0;
    if (g.profiling) { cudaEventRecord(g.end_ev(), g.stream); g.ev_valid = true; }
  }
  CUDA_TRY(cudaGetLastError());
  out->n = (size_t)(M + 1);
  if (location == MRX_OUT_DEVICE) {
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    out->bytes = (uint8_t *)d_bytes;
    out->offsets = (int64_t *)d_offs;
    og.commit();
    return MRX_OK;
  }
  void *hb = nullptr, *ho = nullptr;
  if ((rc = host_alloc((size_t)total + 1, &hb))) return bail(rc);
  own->host[0] = hb;
  if ((rc = host_alloc(sizeof(int64_t) * (M + 2), &ho))) return bail(rc);
  own->host[1] = ho;
  if (total) CUDA_TRY(cudaMemcpyAsync(hb, d_bytes, (size_t)total, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaMemcpyAsync(ho, d_offs, sizeof(int64_t) * (M + 2), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  dev_release(d_bytes); own->dev[0] = nullptr;
  dev_release(d_offs); own->dev[1] = nullptr;
  out->bytes = (uint8_t *)hb;
  out->offsets = (int64_t *)ho;
  og.commit();
  return MRX_OK;
}

}  // extern "C"

// ---- sub / split over a host buffer, pipelined chunk by chunk ----------------------------------------
// A byte that no state of the automaton can consume, that is in neither the matcher class (back-scans) nor
// a candidate of either plan, and that occurs in none of the literals cannot be part of any match.  Cutting
// the text right after such "dead" bytes gives chunks no match can straddle, so sub(text) is the
// concatenation of sub(chunk) and split(text) is the concatenation of the chunks' pieces with the piece
// around every cut glued together.  Chunks are uploaded on the H2D stream, processed on the compute
// stream as they land (mrx_sub / mrx_split on a view of the chunk) and their results travel back on the
// D2H stream: the call costs about max(upload, download) instead of their sum.
namespace {

bool dead_bytes(const Program &p, bool dead[256]) {
  if (p.flag(BF_WILDCARD_ANY) || p.flag(BF_START_ANCHOR) || p.flag(BF_END_ANCHOR) || p.start_accepting()) return false;
  if (p.findall.unsupported || p.search.unsupported) return false;
  bool any = false;
  for (int b = 0; b < 256; ++b) {
    bool d = !p.lookup[b] && !p.findall.cand[b] && !p.search.cand[b];
    for (int q = 0; d && q < p.n_states; ++q) d = p.next(q, b) == kDead;
    for (const std::string *lit : {&p.literal, &p.exact, &p.prefilter, &p.findall.literal, &p.search.literal})
      if (d && lit->find((char)b) != std::string::npos) d = false;
    dead[b] = d;
    any = any || d;
  }
  return any;
}

// chunk starts: 0 = c[0] < c[1] < ... < c.back() = n; every inner cut follows a dead byte
// (sub / split run five kernels and three host synchronisations per chunk: larger chunks than findall's)
constexpr int64_t kEditChunk = 128ll << 20;
std::vector<int64_t> chunk_cuts(const uint8_t *host, int64_t n, const bool dead[256]) {
  std::vector<int64_t> c{0};
  constexpr int64_t kSearch = 1ll << 20;
  while (true) {
    const int64_t nominal = c.back() + kEditChunk;
    if (nominal + kEditChunk / 2 >= n) break;  // the last chunk takes a tail shorter than half a chunk as well
    int64_t cut = -1;
    for (int64_t q = nominal; q < std::min(n, nominal + kSearch); ++q)
      if (dead[host[q]]) { cut = q + 1; break; }
    if (cut < 0 || cut >= n) break;
    c.push_back(cut);
  }
  c.push_back(n);
  return c;
}

struct ChunkedUpload {
  DevTmp buf;
  std::vector<int64_t> dev_off;  // 16-byte aligned device offset of every chunk
  int start(const uint8_t *host, const std::vector<int64_t> &cuts) {
    const size_t k = cuts.size() - 1;
    int rc = buf.get((size_t)cuts.back() + 16 * k + 512);
    if (rc) return rc;
    int64_t a = 0;
    for (size_t i = 0; i < k; ++i) {
      const int64_t len = cuts[i + 1] - cuts[i];
      dev_off.push_back(a);
      CUDA_TRY(cudaMemcpyAsync(buf.as<uint8_t>() + a, host + cuts[i], (size_t)len, cudaMemcpyHostToDevice, g.s_in));
      CUDA_TRY(cudaEventRecord(pipe_event(i), g.s_in));
      a = (a + len + 15) & ~15ll;
    }
    CUDA_TRY(cudaMemsetAsync(buf.as<uint8_t>() + a, 0, 512, g.s_in));
    return MRX_OK;
  }
};

// pinned result buffer that can grow (copies what has been received so far)
struct GrowHost {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t need, size_t used) {
    if (need <= cap) return MRX_OK;
    CUDA_TRY(cudaStreamSynchronize(g.s_out));
    void *q = nullptr;
    const size_t ncap = need + need / 4 + 4096;
    int rc = host_alloc(ncap, &q);
    if (rc) return rc;
    if (p && used) std::memcpy(q, p, used);
    host_release(p);
    p = q;
    cap = ncap;
    return MRX_OK;
  }
};

}  // namespace

// ---- count exchange over NCCL, inside the library ---------------------------------------------------
// The only collective on the path: every rank contributes a small int64 record (match count,
// output bytes) and gets the table of all ranks, from which it derives the global offset of its slice.
// libnccl is not linked: it is looked up in the process (torch has it loaded) or on the default search
// path, so single-GPU users never need it.  The record goes through a ring of pinned slots and the
// all-gather runs on its own stream: a call costs a few microseconds of host time and never waits.
namespace {

struct NcclUid { char internal[128]; };
typedef void *NcclComm;
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(NcclUid *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUid, int) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
constexpr int kXRing = 64, kXWidth = 4, kNcclInt64 = 4;
struct Exchange {
  NcclApi api;
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr, ev_in = nullptr;
  int64_t *h_ring = nullptr;  // pinned, kXRing x kXWidth
  int64_t *d_ring = nullptr;  // device mirror of the slots
  int64_t *d_table = nullptr; // world x width of the last exchange
  uint64_t count = 0;
} gx;

int nccl_load() {
  if (gx.api.h) return MRX_OK;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!h) return fail(MRX_E_UNSUPPORTED, std::string("libnccl.so.2 not found: ") + dlerror());
  NcclApi a;
  a.h = h;
  a.GetUniqueId = (int (*)(NcclUid *))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(NcclComm *, int, NcclUid, int))dlsym(h, "ncclCommInitRank");
  a.AllGather = (int (*)(const void *, void *, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
  a.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
  a.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy) return fail(MRX_E_UNSUPPORTED, "libnccl.so.2 lacks an entry point");
  gx.api = a;
  return MRX_OK;
}
int nccl_fail(const char *what, int code) {
  return fail(MRX_E_CUDA, std::string(what) + ": " + (gx.api.GetErrorString ? gx.api.GetErrorString(code) : "nccl error"));
}

}  // namespace

extern "C" int mrx_comm_unique_id(uint8_t out[128]) {
  int rc = nccl_load();
  if (rc) return rc;
  NcclUid id;
  const int e = gx.api.GetUniqueId(&id);
  if (e) return nccl_fail("ncclGetUniqueId", e);
  std::memcpy(out, id.internal, 128);
  return MRX_OK;
}

extern "C" int mrx_comm_init(const uint8_t id_bytes[128], int rank, int world) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = nccl_load())) return rc;
  if (gx.comm) return MRX_OK;
  if (world < 1 || rank < 0 || rank >= world) return fail(MRX_E_INVALID, "bad rank / world");
  NcclUid id;
  std::memcpy(id.internal, id_bytes, 128);
  const int e = gx.api.CommInitRank(&gx.comm, world, id, rank);
  if (e) { gx.comm = nullptr; return nccl_fail("ncclCommInitRank", e); }
  gx.rank = rank;
  gx.world = world;
  CUDA_TRY(cudaStreamCreateWithFlags(&gx.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&gx.ev, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&gx.ev_in, cudaEventDisableTiming));
  CUDA_TRY(cudaHostAlloc((void **)&gx.h_ring, sizeof(int64_t) * kXRing * kXWidth, cudaHostAllocDefault));
  CUDA_TRY(cudaMalloc((void **)&gx.d_ring, sizeof(int64_t) * kXRing * kXWidth));
  CUDA_TRY(cudaMalloc((void **)&gx.d_table, sizeof(int64_t) * (size_t)world * kXWidth));
  CUDA_TRY(cudaMemset(gx.d_table, 0, sizeof(int64_t) * (size_t)world * kXWidth));
  return MRX_OK;
}

extern "C" int mrx_comm_allgather(const int64_t *values, int width, void **table_dev) {
  if (!gx.comm) return fail(MRX_E_NO_DEVICE, "mrx_comm_init has not been called");
  if (!values || width < 1 || width > kXWidth) return fail(MRX_E_INVALID, "bad record");
  if (gx.count && gx.count % kXRing == 0) CUDA_TRY(cudaStreamSynchronize(gx.stream));  // the ring wraps: slots are free again
  const size_t slot = (size_t)(gx.count++ % kXRing) * kXWidth;
  for (int i = 0; i < width; ++i) gx.h_ring[slot + i] = values[i];
  CUDA_TRY(cudaMemcpyAsync(gx.d_ring + slot, gx.h_ring + slot, sizeof(int64_t) * width, cudaMemcpyHostToDevice, gx.stream));
  const int e = gx.api.AllGather(gx.d_ring + slot, gx.d_table, (size_t)width, kNcclInt64, gx.comm, gx.stream);
  if (e) return nccl_fail("ncclAllGather", e);
  if (table_dev) *table_dev = gx.d_table;
  return MRX_OK;
}

// The same exchange for ONE int64 that is still being produced on the library stream (the count of an asynchronous
// findall, mrx_spans_count_device): the exchange stream waits for the library stream as of now and gathers straight
// from device memory -- the host never sees the value.
extern "C" int mrx_comm_allgather_device(const void *value_dev, void **table_dev) {
  if (!gx.comm) return fail(MRX_E_NO_DEVICE, "mrx_comm_init has not been called");
  if (!value_dev) return fail(MRX_E_INVALID, "null argument");
  CUDA_TRY(cudaEventRecord(gx.ev_in, g.stream));
  CUDA_TRY(cudaStreamWaitEvent(gx.stream, gx.ev_in, 0));
  const int e = gx.api.AllGather(value_dev, gx.d_table, (size_t)1, kNcclInt64, gx.comm, gx.stream);
  if (e) return nccl_fail("ncclAllGather", e);
  if (table_dev) *table_dev = gx.d_table;
  return MRX_OK;
}

// Make `stream` (a cudaStream_t; NULL = the library stream) wait for every exchange issued so far.
extern "C" int mrx_comm_join(void *stream) {
  if (!gx.comm) return MRX_OK;
  CUDA_TRY(cudaEventRecord(gx.ev, gx.stream));
  CUDA_TRY(cudaStreamWaitEvent(stream ? (cudaStream_t)stream : g.stream, gx.ev, 0));
  return MRX_OK;
}

extern "C" int mrx_comm_shutdown(void) {
  if (!gx.comm) return MRX_OK;
  cudaStreamSynchronize(gx.stream);
  gx.api.CommDestroy(gx.comm);
  cudaFreeHost(gx.h_ring);
  cudaFree(gx.d_ring);
  cudaFree(gx.d_table);
  cudaEventDestroy(gx.ev);
  cudaEventDestroy(gx.ev_in);
  cudaStreamDestroy(gx.stream);
  NcclApi keep = gx.api;
  gx = Exchange();
  gx.api = keep;
  return MRX_OK;
}

static int search_batch_host_single(const mrx_regex *re_c, const uint8_t *host, size_t n, const int64_t *doc_offsets,
                                    size_t n_docs, int64_t doc_stride, int64_t *out_start_end) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !out_start_end || (!host && n) || (!doc_offsets && doc_stride <= 0)) return fail(MRX_E_INVALID, "bad argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  if ((rc = ensure_uploaded(re))) return rc;
  if (n_docs == 0) return MRX_OK;
  if (doc_offsets) {
    if (doc_offsets[0] < 0 || doc_offsets[n_docs] > (int64_t)n) return fail(MRX_E_INVALID, "doc_offsets outside the text");
    for (size_t i = 0; i < n_docs; ++i)
      if (doc_offsets[i] > doc_offsets[i + 1]) return fail(MRX_E_INVALID, "doc_offsets must be non-decreasing");
  } else if ((int64_t)(n_docs - 1) > (int64_t)n / doc_stride) {
    return fail(MRX_E_INVALID, "more fixed-size documents than the text holds");
  }
  // documents are independent: groups of documents of about kEditChunk bytes are uploaded, searched and
  // their results downloaded as a three-stream pipeline
  auto doc_begin = [&](size_t d) -> int64_t {
    return doc_offsets ? doc_offsets[d] : std::min<int64_t>((int64_t)d * doc_stride, (int64_t)n);
  };
  std::vector<size_t> first{0};
  while (first.back() < n_docs) {
    size_t d = first.back();
    const int64_t b0 = doc_begin(d);
    size_t hi = n_docs;
    if (doc_offsets) {  // first document that starts kEditChunk bytes further on
      hi = std::upper_bound(doc_offsets + d, doc_offsets + n_docs, b0 + kEditChunk) - doc_offsets;
      if (hi <= d) hi = d + 1;
    } else {
      hi = std::min(n_docs, d + (size_t)std::max<int64_t>(1, kEditChunk / doc_stride));
    }
    first.push_back(hi);
  }
  const size_t k = first.size() - 1;
  DevTmp d_text, d_out, d_offs;
  DrainSideStreams drain;
  if ((rc = d_text.get(n + 512))) return rc;
  if ((rc = d_out.get(sizeof(int64_t) * 2 * n_docs))) return rc;
  if (doc_offsets && (rc = d_offs.get(sizeof(int64_t) * (n_docs + k + 1)))) return rc;
  std::vector<int64_t> rebased;  // per group: offsets relative to the group's first byte (kept alive for the copies)
  if (doc_offsets) rebased.resize(n_docs + k + 1);
  uint8_t *dt = d_text.as<uint8_t>();
  for (size_t i = 0; i < k; ++i) {
    const int64_t b0 = doc_begin(first[i]), b1 = first[i + 1] < n_docs ? doc_begin(first[i + 1]) : (doc_offsets ? doc_offsets[n_docs] : (int64_t)n);
    if (b1 > b0) CUDA_TRY(cudaMemcpyAsync(dt + b0, host + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, g.s_in));
    if (doc_offsets) {
      const size_t base = first[i] + i;
      for (size_t d = first[i]; d <= first[i + 1]; ++d) rebased[base + (d - first[i])] = doc_offsets[d] - b0;
      CUDA_TRY(cudaMemcpyAsync(d_offs.as<int64_t>() + base, rebased.data() + base, sizeof(int64_t) * (first[i + 1] - first[i] + 1),
                               cudaMemcpyHostToDevice, g.s_in));
    }
    CUDA_TRY(cudaEventRecord(pipe_event(i), g.s_in));
  }
  for (size_t i = 0; i < k; ++i) {
    const int64_t b0 = doc_begin(first[i]), b1 = first[i + 1] < n_docs ? doc_begin(first[i + 1]) : (doc_offsets ? doc_offsets[n_docs] : (int64_t)n);
    const size_t nd = first[i + 1] - first[i];
    CUDA_TRY(cudaStreamWaitEvent(g.stream, pipe_event(i), 0));
    mrx_text view;
    view.d = dt + b0; view.n = (size_t)(b1 - b0); view.owned = false; view.core_begin = 0; view.core_end = view.n;
    rc = mrx_search_batch(re_c, &view, doc_offsets ? d_offs.as<int64_t>() + first[i] + i : nullptr, nd, doc_stride, MRX_OUT_DEVICE,
                          d_out.as<int64_t>() + 2 * first[i]);
    if (rc) { cudaStreamSynchronize(g.s_out); cudaStreamSynchronize(g.s_in); cudaStreamSynchronize(g.stream); return rc; }
    CUDA_TRY(cudaEventRecord(pipe_event(k + i), g.stream));
    CUDA_TRY(cudaStreamWaitEvent(g.s_out, pipe_event(k + i), 0));
    CUDA_TRY(cudaMemcpyAsync(out_start_end + 2 * first[i], d_out.as<int64_t>() + 2 * first[i], sizeof(int64_t) * 2 * nd,
                             cudaMemcpyDeviceToHost, g.s_out));
  }
  CUDA_TRY(cudaStreamSynchronize(g.s_out));
  CUDA_TRY(cudaStreamSynchronize(g.s_in));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return MRX_OK;
}

static int sub_host_single(const mrx_regex *re_c, const uint8_t *repl, size_t repl_n, int64_t count, const uint8_t *host,
                           size_t n, mrx_bytes *out) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !out || (!host && n) || (!repl && repl_n)) return fail(MRX_E_INVALID, "null argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  std::memset(out, 0, sizeof *out);
  bool dead[256];
  std::vector<int64_t> cuts;
  if (count == 0 && (int64_t)n >= 2 * kEditChunk && dead_bytes(re->prog, dead)) cuts = chunk_cuts(host, (int64_t)n, dead);
  if (cuts.size() < 3) {  // small text, a replacement limit, or no cut points: upload, then one call
    mrx_text *t = nullptr;
    if ((rc = mrx_text_upload(host, n, &t))) return rc;
    rc = mrx_sub(re_c, repl, repl_n, count, t, MRX_OUT_HOST, out);
    mrx_text_free(t);
    return rc;
  }
  if ((rc = ensure_uploaded(re))) return rc;
  ChunkedUpload up;
  DrainSideStreams drain;  // destroyed before `up`: the side streams are idle when its buffer goes back to the pool
  if ((rc = up.start(host, cuts))) return rc;
  const size_t k = cuts.size() - 1;
  std::vector<mrx_bytes> parts(k);
  GrowHost res;
  size_t off = 0;
  auto cleanup = [&](size_t upto) { for (size_t i = 0; i < upto; ++i) mrx_bytes_free(&parts[i]); };
  for (size_t i = 0; i < k; ++i) {
    const int64_t len = cuts[i + 1] - cuts[i];
    CUDA_TRY(cudaStreamWaitEvent(g.stream, pipe_event(i), 0));
    mrx_text view;
    view.d = up.buf.as<uint8_t>() + up.dev_off[i]; view.n = (size_t)len; view.owned = false; view.core_begin = 0; view.core_end = (size_t)len;
    rc = mrx_sub(re_c, repl, repl_n, 0, &view, MRX_OUT_DEVICE, &parts[i]);  // returns with the chunk's output complete
    if (rc) { cudaStreamSynchronize(g.s_out); cudaStreamSynchronize(g.s_in); cleanup(i); host_release(res.p); return rc; }
    // size the result from the expansion seen so far
    const double ratio = (double)(off + parts[i].n) / (double)cuts[i + 1];
    const size_t guess = (size_t)((double)n * ratio * 1.02) + (1u << 20);
    if ((rc = res.reserve(std::max(off + parts[i].n, i + 1 < k ? guess : off + parts[i].n), off))) { cleanup(i + 1); return rc; }
    if (parts[i].n) CUDA_TRY(cudaMemcpyAsync((uint8_t *)res.p + off, parts[i].bytes, parts[i].n, cudaMemcpyDeviceToHost, g.s_out));
    off += parts[i].n;
  }
  CUDA_TRY(cudaStreamSynchronize(g.s_out));
  CUDA_TRY(cudaStreamSynchronize(g.s_in));
  cleanup(k);
  OwnedOut *own = new OwnedOut();
  own->host[0] = res.p;
  out->_owner = own;
  out->bytes = (uint8_t *)res.p;
  out->n = off;
  out->location = MRX_OUT_HOST;
  return MRX_OK;
}

static int split_host_single(const mrx_regex *re_c, const uint8_t *host, size_t n, int64_t maxsplit, mrx_pieces *out) {
  int rc = require_device();
  if (rc) return rc;
  if (!re_c || !out || (!host && n)) return fail(MRX_E_INVALID, "null argument");
  mrx_regex *re = const_cast<mrx_regex *>(re_c);
  std::memset(out, 0, sizeof *out);
  bool dead[256];
  std::vector<int64_t> cuts;
  if (maxsplit == 0 && (int64_t)n >= 2 * kEditChunk && dead_bytes(re->prog, dead)) cuts = chunk_cuts(host, (int64_t)n, dead);
  if (cuts.size() < 3) {
    mrx_text *t = nullptr;
    if ((rc = mrx_text_upload(host, n, &t))) return rc;
    rc = mrx_split(re_c, t, maxsplit, MRX_OUT_HOST, out);
    mrx_text_free(t);
    return rc;
  }
  if ((rc = ensure_uploaded(re))) return rc;
  ChunkedUpload up;
  DrainSideStreams drain;
  if ((rc = up.start(host, cuts))) return rc;
  const size_t k = cuts.size() - 1;
  std::vector<mrx_pieces> parts(k);
  GrowHost bytes, offs;
  size_t nbytes = 0;    // output bytes received so far
  size_t nsep = 0;      // separators (= inner piece boundaries) so far
  auto cleanup = [&](size_t upto) { for (size_t i = 0; i < upto; ++i) mrx_pieces_free(&parts[i]); };
  auto bail = [&](int code, size_t upto) {
    cudaStreamSynchronize(g.s_out); cudaStreamSynchronize(g.s_in);
    cleanup(upto); host_release(bytes.p); host_release(offs.p);
    return code;
  };
  for (size_t i = 0; i < k; ++i) {
    const int64_t len = cuts[i + 1] - cuts[i];
    CUDA_TRY(cudaStreamWaitEvent(g.stream, pipe_event(i), 0));
    mrx_text view;
    view.d = up.buf.as<uint8_t>() + up.dev_off[i]; view.n = (size_t)len; view.owned = false; view.core_begin = 0; view.core_end = (size_t)len;
    rc = mrx_split(re_c, &view, 0, MRX_OUT_DEVICE, &parts[i]);
    if (rc) return bail(rc, i);
    const size_t m = parts[i].n - 1;  // separators in this chunk
    // bytes of this chunk's pieces = its last offset; then shift its inner boundaries into the whole output
    CUDA_TRY(cudaMemcpyAsync(g.h_scalars + 10, parts[i].offsets + (m + 1), 8, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    const size_t cb = (size_t)g.h_scalars[10];
    launch_add_i64(parts[i].offsets + 1, (int64_t)m, (int64_t)nbytes, g.stream);
    ++g.launches;
    CUDA_TRY(cudaEventRecord(pipe_event(k + i), g.stream));
    const double done = (double)cuts[i + 1];
    const size_t gb = (size_t)((double)n * ((double)(nbytes + cb) / done) * 1.02) + (1u << 20);
    const size_t go = (size_t)((double)n * ((double)(nsep + m) / done) * 1.02) + 4096;
    const bool last = i + 1 == k;
    if ((rc = bytes.reserve(std::max(nbytes + cb, last ? nbytes + cb : gb), nbytes))) return bail(rc, i + 1);
    if ((rc = offs.reserve(sizeof(int64_t) * (std::max(nsep + m, last ? nsep + m : go) + 2), sizeof(int64_t) * (nsep + 1)))) return bail(rc, i + 1);
    CUDA_TRY(cudaStreamWaitEvent(g.s_out, pipe_event(k + i), 0));
    if (cb) CUDA_TRY(cudaMemcpyAsync((uint8_t *)bytes.p + nbytes, parts[i].bytes, cb, cudaMemcpyDeviceToHost, g.s_out));
    if (m) CUDA_TRY(cudaMemcpyAsync((int64_t *)offs.p + 1 + nsep, parts[i].offsets + 1, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, g.s_out));
    nbytes += cb;
    nsep += m;
  }
  CUDA_TRY(cudaStreamSynchronize(g.s_out));
  CUDA_TRY(cudaStreamSynchronize(g.s_in));
  cleanup(k);
  ((int64_t *)offs.p)[0] = 0;
  ((int64_t *)offs.p)[nsep + 1] = (int64_t)nbytes;
  OwnedOut *own = new OwnedOut();
  own->host[0] = bytes.p;
  own->host[1] = offs.p;
  out->_owner = own;
  out->bytes = (uint8_t *)bytes.p;
  out->offsets = (int64_t *)offs.p;
  out->n = nsep + 1;
  out->location = MRX_OUT_HOST;
  return MRX_OK;
}


// ---- multi-device mode: one C-ABI call, all the GPUs of the box ---------------------------------------------------
// mrx_init_multi(n) initialises devices 0 .. n-1 in this process.  The host-buffer entry points then cut their input
// into n contiguous shards -- byte ranges with a halo (findall), document ranges (batch search), ranges that end
// after a byte no match can contain (sub, split) -- and run the single-device implementation on each from its own
// worker thread (own device, own streams, own pools; the threads share nothing but the read-only regex).  No data
// crosses between GPUs: what has to be combined is counts and a few boundary records, and the calling thread is
// their one consumer, so they are combined on the host (the NCCL count exchange, mrx_comm_*, is for the
// one-process-per-GPU layout where no single caller sees all the counts).
namespace {

template <class F>
int run_on_devices(int n, F &&f) {  // f(d) runs with device d current in its own thread; first error wins
  std::vector<int> rcs((size_t)n, MRX_OK);
  std::vector<std::string> errs((size_t)n);
  std::vector<std::thread> threads;
  for (int d = 0; d < n; ++d)
    threads.emplace_back([&, d]() {
      g_cur = &g_devs[d];
      cudaError_t e = cudaSetDevice(d);
      rcs[(size_t)d] = e == cudaSuccess ? f(d) : fail(MRX_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
      errs[(size_t)d] = g_err;
    });
  for (std::thread &t : threads) t.join();
  for (int d = 0; d < n; ++d)
    if (rcs[(size_t)d] != MRX_OK) return fail(rcs[(size_t)d], "device " + std::to_string(d) + ": " + errs[(size_t)d]);
  return MRX_OK;
}

constexpr int64_t kMultiHalo = 1 << 16;

}  // namespace

// byte ranges compose into the whole-text result when no match decision crosses a shard edge: literal sets
// (every occurrence is a match, bounded length) and class runs (clipped at the edges, stitched below)
static bool findall_shardable(const mrx_regex *re) {
  const Plan &pl = re->prog.findall;
  if (pl.unsupported) return false;
  return (pl.dense.valid && pl.direct && pl.dense.max_len < kMultiHalo) || pl.mode == MRX_MODE_CLASS_RUNS;
}

static int findall_host_multi(const mrx_regex *re_c, const uint8_t *host, size_t n, mrx_spans *out) {
  int rc = require_device();
  if (rc) return rc;
  if (!out || (!host && n)) return fail(MRX_E_INVALID, "bad argument");
  std::memset(out, 0, sizeof *out);
  const int N = g_multi;
  const bool runs = re_c->prog.findall.mode == MRX_MODE_CLASS_RUNS;
  std::vector<mrx_spans> parts((size_t)N);
  std::vector<int64_t> lo((size_t)N + 1);
  for (int d = 0; d <= N; ++d) lo[(size_t)d] = (int64_t)((unsigned __int128)n * d / N);
  rc = run_on_devices(N, [&](int d) {
    const int64_t b0 = std::max<int64_t>(0, lo[(size_t)d] - kMultiHalo), b1 = std::min<int64_t>((int64_t)n, lo[(size_t)d + 1] + kMultiHalo);
    return findall_host_shard_single(re_c, host + b0, (size_t)(b1 - b0), (size_t)(lo[(size_t)d] - b0), (size_t)(lo[(size_t)d + 1] - b0), b0,
                                     &parts[(size_t)d]);
  });
  auto free_parts = [&]() { for (mrx_spans &p : parts) mrx_spans_free(&p); };
  if (rc) { free_parts(); return rc; }
  // class runs are clipped to the shard cores: a run that ends on a shard's end and one that starts on the next
  // shard's begin are one run of the reference (it may swallow whole shards)
  std::vector<int64_t> first((size_t)N, 0), count((size_t)N);  // slice [first, first + count) of every part survives
  for (int d = 0; d < N; ++d) count[(size_t)d] = (int64_t)parts[(size_t)d].n;
  if (runs) {
    int d = 0;
    while (d < N) {
      mrx_spans &p = parts[(size_t)d];
      if (count[(size_t)d] == 0 || p.start_end[2 * (first[(size_t)d] + count[(size_t)d] - 1) + 1] != lo[(size_t)d + 1]) { ++d; continue; }
      int64_t *last_end = &p.start_end[2 * (first[(size_t)d] + count[(size_t)d] - 1) + 1];
      int k = d + 1;
      while (k < N) {
        mrx_spans &q = parts[(size_t)k];
        if (count[(size_t)k] == 0 || q.start_end[2 * first[(size_t)k]] != *last_end) break;
        *last_end = q.start_end[2 * first[(size_t)k] + 1];  // the continuation joins the open run ...
        ++first[(size_t)k];                                 // ... and leaves its own shard's list
        --count[(size_t)k];
        if (count[(size_t)k] > 0 || *last_end != lo[(size_t)k + 1]) break;  // the run ended inside shard k
        ++k;                                                // shard k was one run from edge to edge: go on
      }
      d = std::max(k, d + 1);
    }
  }
  int64_t total = 0;
  std::vector<int64_t> at((size_t)N);
  for (int d = 0; d < N; ++d) { at[(size_t)d] = total; total += count[(size_t)d]; }
  void *h = nullptr;
  if ((rc = host_alloc(sizeof(int64_t) * 2 * (size_t)std::max<int64_t>(total, 1), &h))) { free_parts(); return rc; }
  {  // every worker copies its slice into place (the spans are a third of the text's size: not a job for one core)
    std::vector<std::thread> threads;
    for (int d = 0; d < N; ++d)
      threads.emplace_back([&, d]() {
        if (count[(size_t)d] > 0)
          std::memcpy((int64_t *)h + 2 * at[(size_t)d], parts[(size_t)d].start_end + 2 * first[(size_t)d], sizeof(int64_t) * 2 * (size_t)count[(size_t)d]);
      });
    for (std::thread &t : threads) t.join();
  }
  free_parts();
  OwnedOut *own = new OwnedOut();
  own->host[0] = h;
  out->_owner = own;
  out->start_end = (int64_t *)h;
  out->n = (size_t)total;
  out->location = MRX_OUT_HOST;
  return MRX_OK;
}

extern "C" int mrx_search_batch_host(const mrx_regex *re_c, const uint8_t *host, size_t n, const int64_t *doc_offsets,
                                     size_t n_docs, int64_t doc_stride, int64_t *out_start_end) {
  if (g_multi <= 1 || !re_c || !out_start_end || n < (size_t)g_multi * kMultiMinBytes || n_docs < (size_t)g_multi * 1024 ||
      (!doc_offsets && doc_stride <= 0))
    return search_batch_host_single(re_c, host, n, doc_offsets, n_docs, doc_stride, out_start_end);
  int rc = require_device();
  if (rc) return rc;
  const int N = g_multi;
  if (doc_offsets) {
    if (doc_offsets[0] < 0 || doc_offsets[n_docs] > (int64_t)n) return fail(MRX_E_INVALID, "doc_offsets outside the text");
  } else if ((int64_t)(n_docs - 1) > (int64_t)n / doc_stride) {
    return fail(MRX_E_INVALID, "more fixed-size documents than the text holds");
  }
  // documents are independent: contiguous document ranges, results written in place by every device
  return run_on_devices(N, [&](int d) {
    const size_t d0 = n_docs * (size_t)d / N, d1 = n_docs * (size_t)(d + 1) / N;
    if (d1 <= d0) return (int)MRX_OK;
    if (!doc_offsets) {
      const size_t b0 = d0 * (size_t)doc_stride, b1 = std::min(n, d1 * (size_t)doc_stride);
      return search_batch_host_single(re_c, host + b0, b1 - b0, nullptr, d1 - d0, doc_stride, out_start_end + 2 * d0);
    }
    std::vector<int64_t> rebased(d1 - d0 + 1);
    for (size_t i = d0; i <= d1; ++i) rebased[i - d0] = doc_offsets[i] - doc_offsets[d0];
    return search_batch_host_single(re_c, host + doc_offsets[d0], (size_t)(doc_offsets[d1] - doc_offsets[d0]), rebased.data(), d1 - d0,
                                    doc_stride, out_start_end + 2 * d0);
  });
}

// N contiguous parts of the text, every inner cut right after a byte no match can contain; empty when the pattern
// has no such byte or the text none near the nominal cuts
static std::vector<int64_t> device_cuts(const mrx_regex *re, const uint8_t *host, int64_t n, int N) {
  bool dead[256];
  std::vector<int64_t> c{0};
  if (!dead_bytes(re->prog, dead)) return {};
  for (int d = 1; d < N; ++d) {
    const int64_t nominal = (int64_t)((unsigned __int128)n * d / N);
    int64_t cut = -1;
    for (int64_t q = std::max(nominal, c.back()); q < std::min<int64_t>(n, nominal + (4ll << 20)); ++q)
      if (dead[host[q]]) { cut = q + 1; break; }
    if (cut < 0 || cut >= n) return {};
    c.push_back(cut);
  }
  c.push_back(n);
  return c;
}

extern "C" int mrx_sub_host(const mrx_regex *re_c, const uint8_t *repl, size_t repl_n, int64_t count, const uint8_t *host,
                            size_t n, mrx_bytes *out) {
  std::vector<int64_t> cuts;
  if (g_multi > 1 && re_c && out && host && count == 0 && n >= (size_t)g_multi * kMultiMinBytes)
    cuts = device_cuts(re_c, host, (int64_t)n, g_multi);
  if (cuts.empty()) return sub_host_single(re_c, repl, repl_n, count, host, n, out);
  int rc = require_device();
  if (rc) return rc;
  std::memset(out, 0, sizeof *out);
  const int N = g_multi;
  std::vector<mrx_bytes> parts((size_t)N);
  rc = run_on_devices(N, [&](int d) {
    return sub_host_single(re_c, repl, repl_n, 0, host + cuts[(size_t)d], (size_t)(cuts[(size_t)d + 1] - cuts[(size_t)d]), &parts[(size_t)d]);
  });
  auto free_parts = [&]() { for (mrx_bytes &p : parts) mrx_bytes_free(&p); };
  if (rc) { free_parts(); return rc; }
  size_t total = 0;
  std::vector<size_t> at((size_t)N);
  for (int d = 0; d < N; ++d) { at[(size_t)d] = total; total += parts[(size_t)d].n; }
  void *h = nullptr;
  if ((rc = host_alloc(total + 1, &h))) { free_parts(); return rc; }
  {
    std::vector<std::thread> threads;
    for (int d = 0; d < N; ++d)
      threads.emplace_back([&, d]() { if (parts[(size_t)d].n) std::memcpy((uint8_t *)h + at[(size_t)d], parts[(size_t)d].bytes, parts[(size_t)d].n); });
    for (std::thread &t : threads) t.join();
  }
  free_parts();
  OwnedOut *own = new OwnedOut();
  own->host[0] = h;
  out->_owner = own;
  out->bytes = (uint8_t *)h;
  out->n = total;
  out->location = MRX_OUT_HOST;
  return MRX_OK;
}

extern "C" int mrx_split_host(const mrx_regex *re_c, const uint8_t *host, size_t n, int64_t maxsplit, mrx_pieces *out) {
  std::vector<int64_t> cuts;
  if (g_multi > 1 && re_c && out && host && maxsplit == 0 && n >= (size_t)g_multi * kMultiMinBytes)
    cuts = device_cuts(re_c, host, (int64_t)n, g_multi);
  if (cuts.empty()) return split_host_single(re_c, host, n, maxsplit, out);
  int rc = require_device();
  if (rc) return rc;
  std::memset(out, 0, sizeof *out);
  const int N = g_multi;
  std::vector<mrx_pieces> parts((size_t)N);
  rc = run_on_devices(N, [&](int d) {
    return split_host_single(re_c, host + cuts[(size_t)d], (size_t)(cuts[(size_t)d + 1] - cuts[(size_t)d]), 0, &parts[(size_t)d]);
  });
  auto free_parts = [&]() { for (mrx_pieces &p : parts) mrx_pieces_free(&p); };
  if (rc) { free_parts(); return rc; }
  // a cut follows a byte that belongs to a piece: the last piece of a part and the first piece of the next part are
  // one piece of the whole text.  Part d contributes its piece bytes and its INNER piece boundaries.
  size_t nbytes = 0, nsep = 0;
  std::vector<size_t> at_b((size_t)N), at_s((size_t)N);
  for (int d = 0; d < N; ++d) {
    at_b[(size_t)d] = nbytes; at_s[(size_t)d] = nsep;
    nbytes += (size_t)parts[(size_t)d].offsets[parts[(size_t)d].n];
    nsep += parts[(size_t)d].n - 1;
  }
  void *hb = nullptr, *ho = nullptr;
  if ((rc = host_alloc(nbytes + 1, &hb))) { free_parts(); return rc; }
  if ((rc = host_alloc(sizeof(int64_t) * (nsep + 2), &ho))) { host_release(hb); free_parts(); return rc; }
  {
    std::vector<std::thread> threads;
    for (int d = 0; d < N; ++d)
      threads.emplace_back([&, d]() {
        const mrx_pieces &p = parts[(size_t)d];
        const size_t pb = (size_t)p.offsets[p.n];
        if (pb) std::memcpy((uint8_t *)hb + at_b[(size_t)d], p.bytes, pb);
        int64_t *dst = (int64_t *)ho + 1 + at_s[(size_t)d];
        for (size_t i = 1; i < p.n; ++i) dst[i - 1] = p.offsets[i] + (int64_t)at_b[(size_t)d];
      });
    for (std::thread &t : threads) t.join();
  }
  free_parts();
  ((int64_t *)ho)[0] = 0;
  ((int64_t *)ho)[nsep + 1] = (int64_t)nbytes;
  OwnedOut *own = new OwnedOut();
  own->host[0] = hb;
  own->host[1] = ho;
  out->_owner = own;
  out->bytes = (uint8_t *)hb;
  out->offsets = (int64_t *)ho;
  out->n = nsep + 1;
  out->location = MRX_OUT_HOST;
  return MRX_OK;
}

extern "C" {

void mrx_spans_free(mrx_spans *s) {
  if (!s) return;
  if (s->_owner) {
    auto it = async_spans().find(s->_owner);
    if (it != async_spans().end()) {  // still pending: its kernel must be done before the pool hands the blocks on
      if (it->second.done) { cudaEventSynchronize(it->second.done); cudaEventDestroy(it->second.done); }
      async_spans().erase(it);
    }
  }
  release_owner(s->_owner);
  std::memset(s, 0, sizeof *s);
}
void mrx_bytes_free(mrx_bytes *b) {
  if (!b) return;
  release_owner(b->_owner);
  std::memset(b, 0, sizeof *b);
}
void mrx_pieces_free(mrx_pieces *p) {
  if (!p) return;
  release_owner(p->_owner);
  std::memset(p, 0, sizeof *p);
}

}  // extern "C"
