// Packed regex program: the blob format shared by the C++ front end and the Mojo-side
// serializer (mrx_dfa_pack), its decoder, and the planner that picks a scan mode and
// proves (or fails to prove) that potential matches cannot overlap.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "frontend.hpp"
#include "sand.hpp"
#include "sepbits.hpp"

namespace mrx {

constexpr uint32_t kBlobMagic = 0x4258524du;  // "MRXB"
constexpr uint32_t kBlobVersion = 1;
constexpr int kDead = 0xFF;
constexpr int kMaxStates = 254;

enum BlobFlags : uint32_t {
  BF_START_ANCHOR = 1u << 0,
  BF_END_ANCHOR = 1u << 1,
  BF_PURE_LITERAL = 1u << 2,
  BF_HAS_MATCHER = 1u << 3,
  BF_SCAN_ELIGIBLE = 1u << 4,
  BF_WILDCARD_ANY = 1u << 5,
  BF_EXACT_LITERAL = 1u << 6,
  BF_HAS_ANCHORS_INFO = 1u << 7,
  BF_HAS_PREFILTER = 1u << 8,
};

#pragma pack(push, 1)
struct BlobHeader {
  uint32_t magic, version, total_len, flags;
  int32_t n_states, n_classes, start_state, required_byte;
  int32_t fixed_total_width, fixed_num_groups, fixed_concat;
  int32_t fixed_offsets[10], fixed_widths[10];
  uint32_t literal_len, exact_len, prefilter_len, reserved;
  // followed by: byte_class[256] | lookup[256] | accepting[n_states] |
  //              trans[n_states*n_classes] (0xFF = dead) | literal | exact | prefilter
};
#pragma pack(pop)

enum FilterKind : int32_t {
  FILT_NEEDLES = 0,  // candidate byte in a set of <= 4 values: SWAR equality
  FILT_RANGE = 1,    // candidate byte in one contiguous range: SWAR range test
  FILT_LUT = 2,      // 256-entry table in shared memory
  FILT_ALL = 3,      // every position
};

// Dense search automaton for finite, prefix-free literal sets whose occurrences never overlap
// (or a single literal): the Aho-Corasick completion of the reference's trie.  One table lookup per
// text byte, no restarts, no divergence.  The uint16 entry for (row, byte) sits at byte offset
// row + 2 * byte and holds the byte offset of the next row.  Rows are spaced 520 bytes apart (bank
// skew); plain rows start at multiples of 8, "a literal ends here" rows at 8k + 4, so bit 2 of a row
// offset is the accept flag; out_len[j] is the length of the literal ending in accept row j.
struct DenseTable {
  bool valid = false;
  int32_t n_states = 0, acc_base = 0, max_len = 0;
  std::vector<uint16_t> next;    // whole table, table_bytes / 2 entries
  std::vector<uint8_t> out_len;  // n_states - acc_base
  std::vector<uint8_t> len_lut;  // literal length by (row offset >> 3) & 31 of an accept row
};

// Automata of the per-document search kernel (k_lines), both in the layout of DenseTable: uint16 entries
// holding the byte offset of the next row, rows 520 bytes apart (bank skew), flag in bit 2 of a row offset.
//  rev: scanning a document from its last byte to its first, the state after byte i is the set R(i) of
//       forward states from which text[i..] reaches an accepting state (subset construction over those
//       sets).  An attempt of the reference at i succeeds exactly when the start state is in R(i); rows of
//       such sets carry the flag.  Row 0 (offset 0) is the set of accepting states = the state at the
//       document's end.
//  fwd: the reference's automaton itself, one row per state plus a dead row at offset 0; rows of accepting
//       states carry the flag.  fwd_start = row offset of the start state.
struct RevTable {
  bool valid = false;
  int32_t n_rev = 0;
  std::vector<uint16_t> rev, fwd;  // whole tables (bytes / 2 entries each)
  uint16_t fwd_start = 0;
  // a byte every accepted string contains (derived from the forward table, not from the pattern), chosen
  // for rarity, or -1: documents without it cannot match and skip the walk (k_lines_req)
  int32_t req_byte = -1;
};

struct Plan {
  int32_t mode = -1;  // MRX_MODE_*
  DenseTable dense;
  bool sand_valid = false;  // the literal set of `dense` also fits the Shift-And scan (k_sand): findall uses it
  sand::Table sand;
  bool sep_valid = false;   // required-byte plan whose automaton has the shape X* c Y*: split runs fused (k_split)
  sepb::Table sep;
  RevTable rev;
  bool direct = false;        // potential matches provably never overlap: emit them as final
  bool unsupported = false;   // plan cannot be executed faithfully
  std::string why;
  int32_t filt = FILT_LUT;
  int32_t n_needles = 0;
  uint8_t needles[4] = {0, 0, 0, 0};
  uint8_t range_lo = 0, range_hi = 0;
  std::array<uint8_t, 256> cand{};  // candidate predicate K(p) on the byte at p
  bool shortcut = false;            // _try_match_simd fires inside attempts (dfa.mojo:1956-1968)
  bool try_at_end = false;          // loop also attempts p == n (dfa.mojo:2118, 1899)
  int32_t run_min = 0, run_max = 0; // UNIFORM_RUN
  std::string literal;              // LITERAL / EXACT_LITERAL needle
};

struct Program {
  uint32_t flags = 0;
  int32_t n_states = 0, n_classes = 0, start_state = 0, required_byte = -1;
  std::array<uint8_t, 256> byte_class{}, lookup{};
  std::vector<uint8_t> accepting, trans;
  std::string literal, exact, prefilter;
  int32_t fixed_total_width = -1, fixed_num_groups = 0, fixed_concat = 0;
  int32_t fixed_offsets[10] = {0}, fixed_widths[10] = {0};
  std::string blob;
  Plan findall, search;
  // match_next through the MemchrPrefilter (matcher.mojo:784-796, prefilter.mojo:377-430): the search first jumps to
  // the next occurrence of the prefilter literal and runs `search` from there (the engine starts AT the literal:
  // matches that begin before a literal that is not a prefix are missed).  `prefilter_plan`
  // finds the literal's occurrences.  When every match begins with the literal the jump changes nothing and the flag
  // stays down.
  bool search_via_prefilter = false;
  Plan prefilter_plan;

  bool flag(uint32_t f) const { return (flags & f) != 0; }
  bool start_accepting() const { return n_states > 0 && accepting[start_state]; }
  int next(int s, int byte) const { return trans[(size_t)s * n_classes + byte_class[byte]]; }
};

// Reverse automaton of `p` for the attempts of plan `pl`; false when the plan is not an attempt plan
// or the construction exceeds 128 sets.
bool build_reverse(const Program &p, const Plan &pl, RevTable &rt);
// The reference's automaton as a dense table in the RevTable::fwd layout (one row per state plus a dead row
// at offset 0, rows of accepting states flagged in bit 2 of their offset).
void build_forward_table(const Program &p, std::vector<uint16_t> &fwd, uint16_t &start);

// RawDfa (+meta) -> blob.  Returns false (with err) when the automaton does not fit the format.
bool pack_blob(const RawDfa &dfa, const RegexMeta &meta, std::string &blob, std::string &err);
// blob -> Program (validates every length and index); then plans.
bool load_program(const uint8_t *blob, size_t n, Program &p, std::string &err);

}  // namespace mrx
