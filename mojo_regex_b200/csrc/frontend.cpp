// Pattern front end (host only).  See frontend.hpp.  Citations: /root/reference/src/regex/.
#include "frontend.hpp"

#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>

#include "../../include/mojoregex_cuda.h"

namespace mrx {
namespace {

struct ParseError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct UnsupportedError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---- tokens (lexer.mojo:61-195) ----------------------------------------------------------
enum Tok : uint8_t {
  TK_ELEMENT, TK_WILDCARD, TK_SPACE, TK_DIGIT, TK_WORD, TK_START, TK_END, TK_COMMA, TK_LPAREN,
  TK_RPAREN, TK_LCURLY, TK_RCURLY, TK_LBRACKET, TK_RBRACKET, TK_QMARK, TK_STAR, TK_PLUS, TK_VBAR,
  TK_CARET, TK_DASH
};
struct Token {
  Tok type;
  int ch;
  int pos;
};

std::vector<Token> lex(const std::string &re) {
  std::vector<Token> out;
  const int n = (int)re.size();
  bool esc = false;
  for (int i = 0; i < n; ++i) {
    const int c = (uint8_t)re[i];
    if (esc) {
      esc = false;
      switch (c) {
        case 't': out.push_back({TK_ELEMENT, 9, i - 1}); break;  // value slice = '\\' (lexer.mojo:79-82)
        case 's': out.push_back({TK_SPACE, c, i - 1}); break;
        case 'd': out.push_back({TK_DIGIT, c, i - 1}); break;
        case 'w': out.push_back({TK_WORD, c, i - 1}); break;
        default: out.push_back({TK_ELEMENT, c, i}); break;
      }
      continue;
    }
    switch (c) {
      case '\\': esc = true; break;
      case '.': out.push_back({TK_WILDCARD, 0, i}); break;
      case '(': out.push_back({TK_LPAREN, 0, i}); break;
      case ')': out.push_back({TK_RPAREN, 0, i}); break;
      case '[': out.push_back({TK_LBRACKET, 0, i}); break;
      case ']': out.push_back({TK_RBRACKET, 0, i}); break;
      case '-': out.push_back({TK_DASH, 0, i}); break;
      case '{': {  // eager brace scan, lexer.mojo:129-153
        out.push_back({TK_LCURLY, 0, i});
        for (++i; i < n; ++i) {
          const int d = (uint8_t)re[i];
          if (d == ',') out.push_back({TK_COMMA, 0, i});
          else if (d >= '0' && d <= '9') out.push_back({TK_ELEMENT, d, i});
          else if (d == '}') { out.push_back({TK_RCURLY, 0, i}); break; }
          else throw ParseError("Bad token at index " + std::to_string(i));
        }
        break;
      }
      case '^': out.push_back({i == 0 ? TK_START : TK_CARET, 0, i}); break;
      case '$': out.push_back({TK_END, 0, i}); break;
      case '?': out.push_back({TK_QMARK, 0, i}); break;
      case '*': out.push_back({TK_STAR, 0, i}); break;
      case '+': out.push_back({TK_PLUS, 0, i}); break;
      case '|': out.push_back({TK_VBAR, 0, i}); break;
      case '}': out.push_back({TK_RCURLY, 0, i}); break;
      default: out.push_back({TK_ELEMENT, c, i}); break;
    }
  }
  return out;
}

// ---- AST (ast.mojo:25-36, 168-300) --------------------------------------------------------
enum Kind { K_RE, K_ELEMENT, K_WILDCARD, K_SPACE, K_DIGIT, K_WORD, K_RANGE, K_START, K_END, K_OR, K_NOT, K_GROUP };
struct Node {
  int kind;
  int s = 0, e = 0;  // slice of the pattern (value = pattern[s:e], ast.mojo:546-558)
  int mn = 1, mx = 1;
  bool positive = true;
  bool capturing = false;
  int gid = -1;
  std::vector<int> kids;
};
struct Ast {
  std::string pat;
  std::vector<Node> nodes;
  int root = -1;
  int add(Node n) { nodes.push_back(std::move(n)); return (int)nodes.size() - 1; }
  const Node &at(int i) const { return nodes[i]; }
  std::string value(int i) const { const Node &n = nodes[i]; return pat.substr(n.s, n.e - n.s); }
  bool has_value(int i) const { return nodes[i].s != nodes[i].e; }
};

// parser.mojo:56-114
int quantify(const std::vector<Token> &t, int i, Node &el) {
  const int n = (int)t.size();
  const Token &nx = t[i + 1];
  if (nx.type == TK_STAR) { el.mn = 0; el.mx = -1; return i + 1; }
  if (nx.type == TK_PLUS) { el.mn = 1; el.mx = -1; return i + 1; }
  if (nx.type == TK_QMARK) { el.mn = 0; el.mx = 1; return i + 1; }
  if (nx.type != TK_LCURLY) return i;
  i += 2;
  int lo = 0, hi = 0;
  bool has_lo = false, has_hi = false;
  for (; i < n && t[i].type == TK_ELEMENT; ++i) {
    if (t[i].ch < '0' || t[i].ch > '9') throw ParseError("Invalid digit in quantifier");
    lo = lo * 10 + (t[i].ch - '0');
    has_lo = true;
  }
  el.mn = has_lo ? lo : 0;
  if (i < n && t[i].type == TK_COMMA) {
    for (++i; i < n && t[i].type == TK_ELEMENT; ++i) {
      if (t[i].ch < '0' || t[i].ch > '9') throw ParseError("Invalid digit in quantifier");
      hi = hi * 10 + (t[i].ch - '0');
      has_hi = true;
    }
    el.mx = has_hi ? hi : -1;
  } else {
    el.mx = el.mn;
  }
  if (i < n && t[i].type == TK_RCURLY) ++i;
  return i - 1;
}

int make_group(Ast &a, std::vector<int> kids, int s, int e, bool capturing, int gid) {
  Node g;
  g.kind = K_GROUP; g.s = s; g.e = e; g.capturing = capturing; g.gid = gid; g.kids = std::move(kids);
  return a.add(std::move(g));
}

// parser.mojo:130-464
int parse_tokens(Ast &a, const std::vector<Token> &t, int &group_counter) {
  const int n = (int)t.size();
  if (n == 0) return make_group(a, {}, 0, 0, true, 0);
  int depth = 0;
  for (int k = 0; k < n; ++k) {  // top-level '|' -> right-nested binary OR (:150-214)
    if (t[k].type == TK_LPAREN) ++depth;
    else if (t[k].type == TK_RPAREN) --depth;
    else if (t[k].type == TK_VBAR && depth == 0) {
      std::vector<Token> l(t.begin(), t.begin() + k), r(t.begin() + k + 1, t.end());
      int li = l.empty() ? make_group(a, {}, 0, 0, true, 0) : parse_tokens(a, l, group_counter);
      int ri = r.empty() ? make_group(a, {}, 0, 0, true, 0) : parse_tokens(a, r, group_counter);
      Node o;
      o.kind = K_OR; o.s = 0; o.e = (int)a.pat.size(); o.kids = {li, ri};
      return a.add(std::move(o));
    }
  }
  int bd = 0, pd = 0;  // validation (:216-237)
  for (const Token &k : t) {
    if (k.type == TK_LBRACKET) ++bd;
    else if (k.type == TK_RBRACKET) { if (--bd < 0) throw ParseError("Unescaped closing bracket ']' at position " + std::to_string(k.pos)); }
    else if (k.type == TK_LPAREN) ++pd;
    else if (k.type == TK_RPAREN) { if (--pd < 0) throw ParseError("Unescaped closing parenthesis ')' at position " + std::to_string(k.pos)); }
  }
  std::vector<int> elems;
  for (int i = 0; i < n; ++i) {
    const Token &k = t[i];
    switch (k.type) {
      case TK_ELEMENT:
      case TK_DASH: {
        Node el; el.kind = K_ELEMENT; el.s = k.pos; el.e = k.pos + 1;
        if (i + 1 < n) i = quantify(t, i, el);
        elems.push_back(a.add(std::move(el)));
        break;
      }
      case TK_WILDCARD: case TK_SPACE: case TK_DIGIT: case TK_WORD: {
        Node el;
        el.kind = k.type == TK_WILDCARD ? K_WILDCARD : k.type == TK_SPACE ? K_SPACE : k.type == TK_DIGIT ? K_DIGIT : K_WORD;
        el.s = k.pos; el.e = k.pos + (k.type == TK_WILDCARD ? 1 : 2);
        if (i + 1 < n) i = quantify(t, i, el);
        elems.push_back(a.add(std::move(el)));
        break;
      }
      case TK_START: { Node el; el.kind = K_START; el.s = k.pos; el.e = k.pos + 1; elems.push_back(a.add(std::move(el))); break; }
      case TK_END: { Node el; el.kind = K_END; el.s = k.pos; el.e = k.pos + 1; elems.push_back(a.add(std::move(el))); break; }
      case TK_LBRACKET: {  // :313-354, value = raw pattern slice
        const int bs = k.pos;
        ++i;
        bool positive = true;
        if (i < n && t[i].type == TK_CARET) { positive = false; ++i; }
        while (i < n && t[i].type != TK_RBRACKET) {
          if (i + 2 < n && t[i + 1].type == TK_DASH && t[i + 2].type == TK_ELEMENT) i += 3;
          else i += 1;
        }
        if (i >= n) throw ParseError("Missing closing ']'.");
        Node el; el.kind = K_RANGE; el.s = bs; el.e = t[i].pos + 1; el.positive = positive;
        if (i + 1 < n) i = quantify(t, i, el);
        elems.push_back(a.add(std::move(el)));
        break;
      }
      case TK_LPAREN: {  // :365-444
        const int ps = k.pos;
        ++i;
        bool capturing = true;
        int content = ps + 1;
        if (i + 1 < n && t[i].type == TK_QMARK && t[i + 1].type == TK_ELEMENT && t[i + 1].ch == ':') {
          capturing = false; i += 2; content = ps + 3;
        }
        std::vector<Token> inner;
        int cnt = 1;
        while (i < n && cnt > 0) {
          if (t[i].type == TK_LPAREN) ++cnt;
          else if (t[i].type == TK_RPAREN) { if (--cnt == 0) break; }
          inner.push_back(t[i]);
          ++i;
        }
        if (cnt > 0) throw ParseError("Missing closing parenthesis ')'.");
        const int pe = t[i].pos;
        int gid = -1;
        if (capturing) gid = ++group_counter;
        int g = parse_tokens(a, inner, group_counter);
        if (a.nodes[g].kind == K_GROUP) {
          Node &gn = a.nodes[g];
          gn.capturing = capturing; gn.gid = gid; gn.s = content; gn.e = pe;
        } else {
          g = make_group(a, {g}, content, pe, capturing, gid);
        }
        if (i + 1 < n) { Node tmp = a.nodes[g]; i = quantify(t, i, tmp); a.nodes[g].mn = tmp.mn; a.nodes[g].mx = tmp.mx; }
        elems.push_back(g);
        break;
      }
      default: break;  // silently skipped, as the reference does
    }
  }
  return make_group(a, std::move(elems), 0, (int)a.pat.size(), true, 0);
}

Ast parse(const std::string &pattern) {  // parser.mojo:467-510
  Ast a;
  a.pat = pattern;
  int gc = 0;
  int body = parse_tokens(a, lex(pattern), gc);
  Node re; re.kind = K_RE; re.s = 0; re.e = (int)pattern.size(); re.mn = 0; re.mx = 0; re.kids = {body};
  a.root = a.add(std::move(re));
  return a;
}

// ---- classifier (optimizer.mojo:103-846) ---------------------------------------------------
enum { C_SIMPLE = 0, C_MEDIUM = 1, C_COMPLEX = 2 };
constexpr int kMaxLiteralQuant = 10;  // optimizer.mojo:33

int classify_quant(const Node &n) {  // :320-349
  if (n.mn == 1 && n.mx == 1) return C_SIMPLE;
  if (n.mx == -1 && (n.mn == 0 || n.mn == 1)) return C_SIMPLE;
  if (n.mn == 0 && n.mx == 1) return C_SIMPLE;
  if (n.mx != -1 && n.mx - n.mn <= 10) return C_SIMPLE;
  if (n.mx != -1 && n.mx - n.mn <= 100) return C_MEDIUM;
  return C_COMPLEX;
}
bool is_class_kind(int k) { return k == K_RANGE || k == K_DIGIT || k == K_WORD || k == K_SPACE; }

struct Classifier {
  const Ast &a;
  explicit Classifier(const Ast &ast) : a(ast) {}

  bool opt_multi_class_seq(int id) const {  // :512-555
    const Node &g = a.at(id);
    if (g.kind != K_GROUP || g.kids.size() < 2) return false;
    int cc = 0;
    for (int c : g.kids) {
      const Node &e = a.at(c);
      if (is_class_kind(e.kind) || e.kind == K_WILDCARD) ++cc;
      else if (e.kind == K_ELEMENT && e.mn == 1 && e.mx == 1) {}
      else return false;
    }
    return cc >= 2;
  }
  bool all_literal_branches(int id) const {  // :492-510
    const Node &n = a.at(id);
    if (n.kind == K_GROUP) {
      if (n.kids.size() == 1) return all_literal_branches(n.kids[0]);
      for (int c : n.kids) if (a.at(c).kind != K_ELEMENT) return false;
      return true;
    }
    if (n.kind == K_OR) { for (int c : n.kids) if (!all_literal_branches(c)) return false; return true; }
    return n.kind == K_ELEMENT;
  }
  bool simple_quantified_group(const Node &g) const {  // :557-586
    if (!((g.mn == 0 && g.mx == 1) || (g.mn == 0 && g.mx == -1) || (g.mn == 1 && g.mx == -1))) return false;
    for (int c : g.kids) { const Node &e = a.at(c); if (e.kind != K_ELEMENT || e.mn != 1 || e.mx != 1) return false; }
    return true;
  }
  // :648-673, :743-768, dfa.mojo:3462-3486, 3565-3589
  bool strict_branches(int id, std::vector<std::string> &out) const {
    const Node &n = a.at(id);
    if (n.kind == K_OR) return strict_branches(n.kids[0], out) && strict_branches(n.kids[1], out);
    if (n.kind == K_GROUP) {
      std::string s;
      for (int c : n.kids) { if (a.at(c).kind != K_ELEMENT) return false; s += a.value(c); }
      out.push_back(s);
      return true;
    }
    return false;
  }
  static std::string common_prefix(const std::vector<std::string> &b) {  // :675-710
    if (b.empty()) return "";
    if (b.size() == 1) return b[0];
    size_t m = b[0].size();
    for (auto &s : b) m = std::min(m, s.size());
    std::string p;
    for (size_t i = 0; i < m; ++i) {
      for (size_t k = 1; k < b.size(); ++k) if (b[k][i] != b[0][i]) return p;
      p.push_back(b[0][i]);
    }
    return p;
  }
  bool quantified_alt_group(const Node &g) const {  // :712-741
    if (g.mn == 1 && g.mx == 1) return false;
    if (g.kids.size() != 1 || a.at(g.kids[0]).kind != K_OR) return false;
    std::vector<std::string> b;
    return strict_branches(g.kids[0], b);
  }
  bool group_contains_or(int id) const {  // :608-619
    for (int c : a.at(id).kids) {
      if (a.at(c).kind == K_OR) return true;
      if (a.at(c).kind == K_GROUP && group_contains_or(c)) return true;
    }
    return false;
  }
  bool nested_alternation(int id) const {  // :588-606
    for (int c : a.at(id).kids) {
      if (a.at(c).kind == K_OR) { if (nested_alternation(c)) return true; }
      else if (a.at(c).kind == K_GROUP) { if (group_contains_or(c)) return true; }
    }
    return false;
  }
  bool simple_dfa_node(const Node &n) const {  // :827-845
    if (n.kind == K_ELEMENT || is_class_kind(n.kind) || n.kind == K_WILDCARD) return n.mx <= 10 || n.mx == -1;
    return n.kind == K_START || n.kind == K_END;
  }
  bool dfa_compatible_branch(int id) const {  // :794-825
    const Node &n = a.at(id);
    if (n.kind == K_ELEMENT || is_class_kind(n.kind)) return true;
    if (n.kind == K_GROUP) {
      if (n.kids.size() <= 4) { for (int c : n.kids) if (!simple_dfa_node(a.at(c))) return false; return true; }
    } else if (n.kind == K_OR) {
      if (n.kids.size() <= 4) { for (int c : n.kids) if (!dfa_compatible_branch(c)) return false; return true; }
    }
    return false;
  }
  bool literal_heavy(int id) const {  // :770-792
    const Node &n = a.at(id);
    if (n.kind != K_OR) return false;
    int ok = 0;
    for (int c : n.kids) ok += dfa_compatible_branch(c);
    return ok * 5 >= (int)n.kids.size() * 4;
  }
  bool common_prefix_tree(int id) const {  // :621-646
    std::vector<std::string> b;
    if (!strict_branches(id, b) || b.size() < 2) return false;
    return common_prefix(b).size() >= 2;
  }
  int alternation(int id, int depth) const {  // :351-401
    const Node &n = a.at(id);
    if (depth > 2) {
      if (literal_heavy(id)) return C_MEDIUM;
      if (depth <= 4 && common_prefix_tree(id)) return C_SIMPLE;
      return C_COMPLEX;
    }
    int mx = C_SIMPLE;
    for (int c : n.kids) {
      int bc = node(c, depth + 1);
      if (bc == C_COMPLEX) return C_COMPLEX;
      if (bc == C_MEDIUM) mx = C_MEDIUM;
    }
    if (mx == C_SIMPLE && n.kids.size() <= 8 && !nested_alternation(id)) return C_SIMPLE;
    return C_MEDIUM;
  }
  int group(int id, int depth) const {  // :403-490
    const Node &g = a.at(id);
    if (depth > 4) return C_COMPLEX;
    const int qc = classify_quant(g);
    if (qc == C_COMPLEX) return C_COMPLEX;
    if (g.mn != 1 || g.mx != 1) {
      if (simple_quantified_group(g)) {}
      else if (quantified_alt_group(g)) return C_SIMPLE;
      else return C_MEDIUM;
    }
    if (g.kids.size() == 1) {
      const Node &only = a.at(g.kids[0]);
      if ((only.kind == K_OR || only.kind == K_GROUP) && all_literal_branches(g.kids[0])) return C_SIMPLE;
    }
    int mx = C_SIMPLE;
    for (int c : g.kids) {
      int cc = node(c, depth + 1);
      if (cc == C_COMPLEX) return C_COMPLEX;
      if (cc == C_MEDIUM) mx = C_MEDIUM;
    }
    if (mx == C_SIMPLE && qc == C_SIMPLE) {
      bool all_lit = true;
      for (int c : g.kids) {
        const Node &e = a.at(c);
        bool lit = e.kind == K_ELEMENT && e.mn == e.mx && e.mn >= 1 && e.mn <= kMaxLiteralQuant;
        if (!(lit || e.kind == K_START || e.kind == K_END)) { all_lit = false; break; }
      }
      if (all_lit && g.kids.size() <= 20) return C_SIMPLE;
      if (g.kids.size() <= 5) return C_SIMPLE;
      return C_MEDIUM;
    }
    return C_MEDIUM;
  }
  int node(int id, int depth) const {  // :260-318
    const Node &n = a.at(id);
    switch (n.kind) {
      case K_RE: return n.kids.empty() ? C_SIMPLE : node(n.kids[0], depth);
      case K_ELEMENT: case K_WILDCARD: case K_SPACE: case K_DIGIT: case K_WORD: case K_RANGE: return classify_quant(n);
      case K_START: case K_END: return C_SIMPLE;
      case K_OR: return alternation(id, depth);
      case K_GROUP: return opt_multi_class_seq(id) ? C_SIMPLE : group(id, depth);
      default: return C_COMPLEX;
    }
  }
  int simd_nodes(int id) const {  // :234-258
    const Node &n = a.at(id);
    if (is_class_kind(n.kind)) return (n.mn > 1 || n.mx == -1) ? 2 : 1;
    int c = 0;
    if (n.kind == K_GROUP || n.kind == K_RE || n.kind == K_OR) for (int k : n.kids) c += simd_nodes(k);
    return c;
  }
};

// optimizer.mojo:866-900
bool literal_sequence(const Ast &a, int id) {
  const Node &n = a.at(id);
  if (n.kind == K_ELEMENT) return (n.mn == 1 && n.mx == 1) || (n.mn == n.mx && n.mn >= 1 && n.mn <= kMaxLiteralQuant);
  if (n.kind == K_START || n.kind == K_END) return true;
  if (n.kind == K_GROUP) {
    for (int c : n.kids) { if (a.at(c).kind == K_GROUP) return false; if (!literal_sequence(a, c)) return false; }
    return true;
  }
  return false;
}
bool is_literal_pattern(const Ast &a) {  // :848-863
  const Node &r = a.at(a.root);
  if (r.kids.empty()) return true;
  return literal_sequence(a, r.kids[0]);
}
std::string literal_chars(const Ast &a, int id) {  // :921-953
  const Node &n = a.at(id);
  if (n.kind == K_ELEMENT) {
    if (!a.has_value(id)) return "";
    std::string v = a.value(id), out;
    if (n.mn <= 1) return v;
    for (int i = 0; i < n.mn; ++i) out += v;
    return out;
  }
  if (n.kind == K_GROUP) { std::string out; for (int c : n.kids) out += literal_chars(a, c); return out; }
  return "";
}
void anchors_rec(const Ast &a, int id, bool &s, bool &e) {  // :975-999
  const Node &n = a.at(id);
  if (n.kind == K_START) s = true;
  else if (n.kind == K_END) e = true;
  else if (n.kind == K_GROUP) for (int c : n.kids) anchors_rec(a, c, s, e);
}
void pattern_anchors(const Ast &a, bool &s, bool &e) {  // :956-972
  s = e = false;
  const Node &r = a.at(a.root);
  if (!r.kids.empty()) anchors_rec(a, r.kids[0], s, e);
}

// ---- literal analysis for the prefilter (literal_optimizer.mojo) ------------------------------
struct Lit {
  std::string text;
  int offset;
  bool prefix, required;
};
void extract_sequence(const Ast &a, int gid, int off, bool required, bool at_start, std::vector<Lit> &out) {  // :333-394
  std::string cur;
  int cur_off = off;
  bool seq_start = at_start;
  for (int c : a.at(gid).kids) {
    const Node &ch = a.at(c);
    if (ch.kind == K_ELEMENT && ch.mn == 1 && ch.mx == 1 && a.has_value(c)) { cur += a.value(c); continue; }
    if (!cur.empty()) {
      out.push_back({cur, cur_off, seq_start, required});
      cur_off += (int)cur.size();
      seq_start = false;
      cur.clear();
    }
    if (ch.kind == K_START || ch.kind == K_END) continue;
    seq_start = false;
    if (ch.mn > 0) cur_off += 1;
  }
  if (!cur.empty()) out.push_back({cur, cur_off, seq_start, required});
}
std::string prefix_literal(const Ast &a, int id) {  // :430-447
  const Node &n = a.at(id);
  if (n.kind == K_ELEMENT && n.mn >= 1 && n.mx >= 1 && a.has_value(id)) return a.value(id);
  if (n.kind == K_GROUP && n.mn >= 1) {
    std::vector<Lit> l;
    extract_sequence(a, id, 0, true, true, l);
    if (!l.empty()) return l[0].text;
  }
  return "";
}
void or_prefixes(const Ast &a, int id, std::vector<std::string> &out) {  // :416-427
  if (a.at(id).kind != K_OR) { std::string p = prefix_literal(a, id); if (!p.empty()) out.push_back(p); return; }
  for (int c : a.at(id).kids) or_prefixes(a, c, out);
}
std::string or_common_prefix(const Ast &a, int id) {  // :397-413
  std::vector<std::string> p;
  or_prefixes(a, id, p);
  if (p.size() < 2) return "";
  std::string common = p[0];
  for (size_t i = 1; i < p.size(); ++i) {
    size_t k = 0;
    while (k < common.size() && k < p[i].size() && common[k] == p[i][k]) ++k;
    common.resize(k);
    if (common.empty()) return "";
  }
  return common;
}
void extract_from(const Ast &a, int id, std::vector<Lit> &out, int off, bool required, bool at_start) {  // :241-330
  const Node &n = a.at(id);
  if (n.kind == K_ELEMENT) {
    if (n.mn >= 1 && a.has_value(id) && (n.mx == 1 || n.mx == -1)) out.push_back({a.value(id), off, at_start, required});
  } else if (n.kind == K_GROUP) {
    if (n.mn >= 1) {
      if (n.kids.size() == 1) {
        int k = a.at(n.kids[0]).kind;
        if (k == K_GROUP || k == K_OR) { extract_from(a, n.kids[0], out, off, required, at_start); return; }
      }
      extract_sequence(a, id, off, required, at_start, out);
    }
  } else if (n.kind == K_OR) {
    std::string cp = or_common_prefix(a, id);
    if (!cp.empty()) out.push_back({cp, off, at_start, true});
    for (int c : n.kids) extract_from(a, c, out, off, false, at_start);
  }
}
bool best_literal(const Ast &a, Lit &best) {  // :217-238 + select_best :166-206
  std::vector<Lit> lits;
  const Node &r = a.at(a.root);
  if (!r.kids.empty()) extract_from(a, r.kids[0], lits, 0, true, true);
  if (lits.empty()) return false;
  int bi = 0, bs = 0;
  for (int i = 0; i < (int)lits.size(); ++i) {
    int sc = (lits[i].required ? 1000 : 0) + (int)lits[i].text.size() * 10 + (lits[i].prefix ? 100 : 0) + lits[i].offset;
    if (sc > bs) { bs = sc; bi = i; }
  }
  best = lits[bi];
  return true;
}
bool literal_prefix_node(const Ast &a, int id) {  // :479-497
  const Node &n = a.at(id);
  if (n.kind == K_ELEMENT) return n.mn >= 1 && n.mx >= 1;
  if (n.kind == K_GROUP && n.mn >= 1 && !n.kids.empty()) {
    if (a.at(n.kids[0]).kind == K_START && n.kids.size() > 1) return literal_prefix_node(a, n.kids[1]);
    return literal_prefix_node(a, n.kids[0]);
  }
  return false;
}
bool ast_has_anchor(const Ast &a, int id) {  // matcher.mojo:168-178
  const Node &n = a.at(id);
  if (n.kind == K_START || n.kind == K_END) return true;
  if (n.kind == K_GROUP || n.kind == K_RE) for (int c : n.kids) if (ast_has_anchor(a, c)) return true;
  return false;
}
bool skip_prefilter(const std::string &p) {  // matcher.mojo:447-532
  const int n = (int)p.size();
  if (n <= 4) return true;
  bool q = false, alt = false, anch = false, wild = false;
  int bars = 0;
  for (int i = 0; i < n; ++i) {
    char c = p[i];
    if (c == '*' || c == '+' || c == '?') q = true;
    else if (c == '|') { alt = true; ++bars; }
    else if (c == '^' && i == 0) anch = true;
    else if (c == '$' && i == n - 1) anch = true;
    else if (c == '.') wild = true;
  }
  if (q && !alt && !wild) return true;
  if (alt && n <= 10 && !wild && bars >= n / 3) return true;
  if (anch && n <= 10) return true;
  if (wild && n >= 8) return false;
  if (alt && n >= 12) return false;
  if (n >= 15) return false;
  return true;
}

// ---- automaton builders (dfa.mojo) -------------------------------------------------------------
const std::string kDigits = "0123456789";
const std::string kLower = "abcdefghijklmnopqrstuvwxyz";
const std::string kUpper = "ABCDEFGHIJKLMNOPQRSTUVWXYZ";
const std::string kWord = kLower + kUpper + kDigits + "_";  // aliases.mojo:7-9
const std::string kSpace = " \t\n\r\f";                      // dfa.mojo:89, 2901-2902
std::string all_except_newline() {                            // aliases.mojo:1-3
  std::string s;
  for (int i = 32; i < 127; ++i) if (i != '\n') s.push_back((char)i);
  return s;
}

struct Builder {
  RawDfa &d;
  explicit Builder(RawDfa &dfa) : d(dfa) {}
  int state(bool acc = false) {
    std::array<int32_t, 256> row;
    row.fill(-1);
    d.trans.push_back(row);
    d.accepting.push_back(acc ? 1 : 0);
    return (int)d.trans.size() - 1;
  }
  void reset() { d.trans.clear(); d.accepting.clear(); }
  void edge(int s, int c, int t) { if (c >= 0 && c < 256) d.trans[s][c] = t; }  // later writes win (:233-242)
  void cls(int from, int to, const std::string &cc, bool positive) {           // :1747-1803
    if (from >= (int)d.trans.size()) return;
    if (positive) { for (unsigned char c : cc) edge(from, c, to); }
    else { d.trans[from].fill(to); for (unsigned char c : cc) d.trans[from][c] = -1; }
  }
  int find_or_create(int from, int c) {  // :1245-1268
    int t = d.trans[from][c];
    if (t != -1) return t;
    int ns = state();
    edge(from, c, ns);
    return ns;
  }
  void matcher(const std::string &cc) {  // get_character_class_matcher, simd_ops.mojo:1177-1216
    d.has_matcher = true;
    d.lookup.fill(0);
    if (cc == " \t\n\r\f" || cc == " \t\n\r\f\v") { for (unsigned char c : std::string(" \t\n\r\f\v")) d.lookup[c] = 1; return; }
    for (unsigned char c : cc) d.lookup[c] = 1;
  }
};

std::string expand_range(const std::string &r) {  // dfa.mojo:71-168 (node type is always RE at the call sites)
  if (r.size() < 2 || r.front() != '[' || r.back() != ']') return r;
  if (r == "[a-z]") return kLower;
  if (r == "[A-Z]") return kUpper;
  if (r == "[0-9]") return kDigits;
  if (r == "[a-zA-Z0-9]") return kLower + kUpper + kDigits;
  if (r == "[a-zA-Z]") return kLower + kUpper;
  std::string in = r.substr(1, r.size() - 2);
  if (!in.empty() && in[0] == '^') in = in.substr(1);
  auto slice = [](const std::string &s, int a, int b) -> std::string {
    a = std::max(a, 0); b = std::min(b, (int)s.size());
    return a < b ? s.substr(a, b - a) : std::string();
  };
  if (in.size() == 3 && in[1] == '-') {
    int s = (uint8_t)in[0], e = (uint8_t)in[2];
    if (s >= 'a' && e <= 'z') return slice(kLower, s - 'a', e - 'a' + 1);
    if (s >= 'a' && e <= 'Z') return slice(kUpper, s - 'A', e - 'A' + 1);
    if (s >= '0' && e <= '9') return slice(kDigits, s - '0', e - '0' + 1);
  }
  std::string out;
  for (size_t i = 0; i < in.size();) {
    if (i + 2 < in.size() && in[i + 1] == '-') {
      for (int c = (uint8_t)in[i]; c <= (uint8_t)in[i + 2]; ++c) out.push_back((char)c);
      i += 3;
    } else {
      out.push_back(in[i]);
      ++i;
    }
  }
  return out;
}

struct SeqEl {  // dfa.mojo:171-196
  std::string cc;
  int mn, mx;
  bool positive = true;
  std::vector<std::string> branches;
};

struct Shapes {
  const Ast &a;
  explicit Shapes(const Ast &ast) : a(ast) {}
  const Node &root() const { return a.at(a.root); }

  std::string element_class(int id) const {  // :2885-2908
    const Node &n = a.at(id);
    switch (n.kind) {
      case K_DIGIT: return kDigits;
      case K_WORD: return kWord;
      case K_RANGE: return expand_range(a.value(id));
      case K_SPACE: return kSpace;
      case K_WILDCARD: return all_except_newline();
      case K_ELEMENT: return a.has_value(id) ? a.value(id) : "";
      default: return "";
    }
  }
  bool char_class_group(int id) const {  // :2711-2733
    const Node &n = a.at(id);
    if (n.kind != K_GROUP) return false;
    bool has = false;
    for (int c : n.kids) { if (element_class(c).empty()) return false; if (a.at(c).kind != K_ELEMENT) has = true; }
    return has;
  }
  bool literal_alt_group(int id) const {  // :2736-2771
    const Node &n = a.at(id);
    if (n.kind != K_GROUP || n.kids.size() != 1 || a.at(n.kids[0]).kind != K_OR) return false;
    bool has = false;
    std::vector<int> st{n.kids[0]};
    while (!st.empty()) {
      int cur = st.back(); st.pop_back();
      const Node &c = a.at(cur);
      if (c.kind == K_GROUP) { for (int k : c.kids) if (a.at(k).kind != K_ELEMENT) return false; has = true; }
      else if (c.kind == K_OR) { if (c.kids.size() != 2) return false; st.push_back(c.kids[1]); st.push_back(c.kids[0]); }
      else return false;
    }
    return has;
  }
  std::vector<std::string> alt_group_branches(int id) const {  // :2774-2812
    std::vector<std::string> out;
    std::vector<int> st;
    if (!a.at(id).kids.empty()) st.push_back(a.at(id).kids[0]);
    while (!st.empty()) {
      int cur = st.back(); st.pop_back();
      const Node &c = a.at(cur);
      if (c.kind == K_GROUP) {
        std::string b;
        for (int k : c.kids) if (a.at(k).kind == K_ELEMENT && a.has_value(k)) b += a.value(k);
        out.push_back(b);
      } else if (c.kind == K_OR) {
        if (c.kids.size() >= 2) st.push_back(c.kids[1]);
        if (c.kids.size() >= 1) st.push_back(c.kids[0]);
      }
    }
    return out;
  }
  int sole_group() const {  // RE -> single GROUP child, else -1
    const Node &r = root();
    if (r.kids.size() != 1 || a.at(r.kids[0]).kind != K_GROUP) return -1;
    return r.kids[0];
  }
  bool multi_class_seq() const {  // :2815-2882
    int g = sole_group();
    if (g < 0 || a.at(g).kids.size() < 2) return false;
    int cc = 0;
    for (int c : a.at(g).kids) {
      const Node &e = a.at(c);
      if (is_class_kind(e.kind) || e.kind == K_WILDCARD) ++cc;
      else if (e.kind == K_ELEMENT && e.mn == 1 && e.mx == 1) {}
      else if (e.kind == K_GROUP && literal_alt_group(c)) {}
      else if (e.kind == K_GROUP && char_class_group(c)) ++cc;
      else return false;
    }
    return cc >= 2;
  }
  std::vector<SeqEl> multi_class_info() const {  // :2911-2970
    std::vector<SeqEl> els;
    int g = sole_group();
    if (g < 0) return els;
    for (int c : a.at(g).kids) {
      const Node &e = a.at(c);
      if (e.kind == K_GROUP && literal_alt_group(c)) {
        SeqEl s{"", 1, 1, true, alt_group_branches(c)};
        els.push_back(s);
      } else if (e.kind == K_GROUP && char_class_group(c)) {
        for (int k : e.kids) { std::string cc = element_class(k); if (!cc.empty()) els.push_back({cc, a.at(k).mn, a.at(k).mx, a.at(k).positive, {}}); }
      } else {
        std::string cc = element_class(c);
        if (!cc.empty()) els.push_back({cc, e.mn, e.mx, e.positive, {}});
      }
    }
    return els;
  }
  bool simple_class() const {  // :2499-2529
    if (multi_class_seq()) return false;
    const Node &r = root();
    if (r.kids.size() != 1) return false;
    const Node &c = a.at(r.kids[0]);
    auto leaf = [](int k) { return k == K_DIGIT || k == K_WORD || k == K_RANGE; };
    if (leaf(c.kind)) return true;
    if (c.kind == K_GROUP && c.kids.size() == 1) return leaf(a.at(c.kids[0]).kind);
    return false;
  }
  bool pure_anchor(int id) const {  // :2605-2629
    const Node &n = a.at(id);
    if (n.kind == K_START || n.kind == K_END) return true;
    if (n.kind == K_RE) return !n.kids.empty() && pure_anchor(n.kids[0]);
    if (n.kind == K_GROUP) { for (int c : n.kids) if (!pure_anchor(c)) return false; return true; }
    return false;
  }
  bool sequential_class() const {  // :2632-2663
    int g = sole_group();
    if (g < 0) return false;
    for (int c : a.at(g).kids) { int k = a.at(c).kind; if (k != K_RANGE && k != K_DIGIT && k != K_WORD) return false; }
    return a.at(g).kids.size() >= 2;
  }
  bool mixed_sequential() const {  // :2973-3015
    int g = sole_group();
    if (g < 0 || a.at(g).kids.size() < 3) return false;
    bool cc = false, opt = false;
    for (int c : a.at(g).kids) {
      const Node &e = a.at(c);
      if (e.kind == K_RANGE || e.kind == K_DIGIT || e.kind == K_WORD) cc = true;
      else if (e.kind == K_ELEMENT && e.mn == 0 && e.mx == 1) opt = true;
    }
    return cc && opt;
  }
  bool only_literals(int id) const {  // :3100-3120
    if (a.at(id).kind != K_GROUP) return false;
    for (int c : a.at(id).kids) if (a.at(c).kind != K_ELEMENT) return false;
    return true;
  }
  int unwrap_single_groups(int id) const {
    while (a.at(id).kind == K_GROUP && a.at(id).kids.size() == 1) id = a.at(id).kids[0];
    return id;
  }
  bool simple_alt_branches(int id) const {  // :3054-3097
    if (a.at(id).kind != K_OR) return false;
    for (int b : a.at(id).kids) {
      const Node &br = a.at(b);
      if (br.kind == K_GROUP) {
        if (!only_literals(b)) {
          int inner = unwrap_single_groups(b);
          if (a.at(inner).kind == K_OR) { if (!simple_alt_branches(inner)) return false; }
          else return false;
        }
      } else if (br.kind == K_ELEMENT) {
      } else if (br.kind == K_OR) { if (!simple_alt_branches(b)) return false; }
      else return false;
    }
    return true;
  }
  bool alternation() const {  // :3034-3051, 3309-3342
    const Node &r = root();
    if (r.kids.size() != 1) return false;
    int c = r.kids[0];
    if (a.at(c).kind == K_OR) return simple_alt_branches(c);
    int n = unwrap_single_groups(c);
    return a.at(n).kind == K_OR && simple_alt_branches(n);
  }
  int find_or(int id) const {  // :3146-3169
    if (a.at(id).kind == K_OR) return id;
    for (int c : a.at(id).kids) { int f = find_or(c); if (f >= 0) return f; }
    return -1;
  }
  std::string branch_text(int id) const {  // :3172-3194
    const Node &n = a.at(id);
    if (n.kind == K_ELEMENT && a.has_value(id)) return a.value(id);
    std::string s;
    if (n.kind == K_GROUP) for (int c : n.kids) if (a.at(c).kind == K_ELEMENT && a.has_value(c)) s += a.value(c);
    return s;
  }
  void all_alt_branches(int or_id, std::vector<std::string> &out) const {  // :3265-3306
    for (int b : a.at(or_id).kids) {
      const Node &br = a.at(b);
      if (br.kind == K_OR) all_alt_branches(b, out);
      else if (br.kind == K_GROUP) {
        int inner = unwrap_single_groups(b);
        if (a.at(inner).kind == K_OR) all_alt_branches(inner, out);
        else { std::string t = branch_text(b); if (!t.empty()) out.push_back(t); }
      } else { std::string t = branch_text(b); if (!t.empty()) out.push_back(t); }
    }
  }
  int quantified_inner_group() const {  // RE -> GROUP(1 child) -> GROUP ; -1 otherwise
    int g = sole_group();
    if (g < 0 || a.at(g).kids.size() != 1) return -1;
    int in = a.at(g).kids[0];
    return a.at(in).kind == K_GROUP ? in : -1;
  }
  bool quantified_group() const {  // :3197-3242
    int in = quantified_inner_group();
    if (in < 0) return false;
    const Node &g = a.at(in);
    if (g.mn == 1 && g.mx == 1) return false;
    for (int c : g.kids) if (a.at(c).kind != K_ELEMENT) return false;
    return true;
  }
  bool simple_quantifier() const {  // :3345-3383
    int g = sole_group();
    if (g < 0 || a.at(g).kids.empty()) return false;
    bool q = false;
    for (int c : a.at(g).kids) {
      const Node &e = a.at(c);
      if (e.kind != K_ELEMENT) return false;
      if ((e.mn == 0 && e.mx == -1) || (e.mn == 1 && e.mx == -1) || (e.mn == 0 && e.mx == 1)) q = true;
    }
    return q;
  }
  bool wildcard_quantifier() const {  // :3386-3415
    int g = sole_group();
    if (g < 0 || a.at(g).kids.size() != 1) return false;
    const Node &w = a.at(a.at(g).kids[0]);
    if (w.kind != K_WILDCARD) return false;
    return (w.mn == 0 && w.mx == -1) || (w.mn == 1 && w.mx == -1) || (w.mn == 0 && w.mx == 1) || (w.mn == 1 && w.mx == 1);
  }
  int double_group_or() const {  // RE -> GROUP -> GROUP -> OR ; -1 otherwise
    int in = quantified_inner_group();
    if (in < 0 || a.at(in).kids.size() != 1) return -1;
    int o = a.at(in).kids[0];
    return a.at(o).kind == K_OR ? o : -1;
  }
  void lenient_branches(int id, std::vector<std::string> &out) const {  // :1396-1418, 1556-1581
    const Node &n = a.at(id);
    if (n.kind == K_OR) { lenient_branches(n.kids[0], out); lenient_branches(n.kids[1], out); }
    else if (n.kind == K_GROUP) {
      std::string s;
      for (int c : n.kids) if (a.at(c).kind == K_ELEMENT) s += a.value(c);
      out.push_back(s);
    }
  }
};

void build_literal(Builder &b, const std::string &lit, bool hs, bool he) {  // :308-352
  RawDfa &d = b.d;
  d.has_start_anchor = hs; d.has_end_anchor = he; d.literal = lit;
  if (lit.empty()) { b.state(true); return; }
  if (!hs && !he) d.is_pure_literal = true;
  for (size_t i = 0; i < lit.size(); ++i) { int s = b.state(); b.edge(s, (uint8_t)lit[i], (int)i + 1); }
  b.state(true);
}

void build_single_class(Builder &b, const std::string &cc, int mn, int mx, bool positive) {  // :375-496
  RawDfa &d = b.d;
  if (mn >= 0 && positive) { b.matcher(cc); d.scan_eligible = (mx == -1); }
  if (mn == 0) {
    b.state(true); b.state(true);
    b.cls(0, 1, cc, positive);
    if (mx == -1 || mx > 1) b.cls(1, 1, cc, positive);
  } else if (mn == 1) {
    b.state(false); b.state(true);
    b.cls(0, 1, cc, positive);
    if (mx == -1) b.cls(1, 1, cc, positive);
    else if (mx > 1) for (int k = 2; k <= mx; ++k) { b.state(true); b.cls(k - 1, k, cc, positive); }
  } else {
    for (int k = 0; k <= mn; ++k) { b.state(k >= mn); if (k > 0) b.cls(k - 1, k, cc, positive); }
    if (mx == -1) { int last = (int)d.trans.size() - 1; b.cls(last, last, cc, positive); }
    else if (mx > mn) for (int k = mn + 1; k <= mx; ++k) { b.state(true); b.cls(k - 1, k, cc, positive); }
  }
}

void build_sequential(Builder &b, const std::vector<SeqEl> &els, bool hs, bool he) {  // :498-605
  RawDfa &d = b.d;
  d.has_start_anchor = hs; d.has_end_anchor = he;
  if (els.empty()) { b.state(true); return; }
  int cur = 0;
  for (size_t i = 0; i < els.size(); ++i) {
    const SeqEl &el = els[i];
    const bool last = i + 1 == els.size();
    if (el.mn == 0) {
      if (i == 0) { b.state(!last); cur = 0; }
      int ms = b.state(true);
      b.cls(cur, ms, el.cc, el.positive);
      if (el.mx == -1) b.cls(ms, ms, el.cc, el.positive);
      cur = ms;
    } else {
      for (int k = 0; k < el.mn; ++k) {
        int si = b.state(k >= el.mn - 1 && last);
        b.cls(k == 0 ? cur : si - 1, si, el.cc, el.positive);
        cur = si;
      }
      if (el.mx == -1) b.cls(cur, cur, el.cc, el.positive);
      else if (el.mx > el.mn) for (int k = 0; k < el.mx - el.mn; ++k) { int si = b.state(last); b.cls(cur, si, el.cc, el.positive); cur = si; }
    }
  }
}

void build_multi_class(Builder &b, const std::vector<SeqEl> &els, bool hs, bool he) {  // :607-871
  RawDfa &d = b.d;
  d.has_start_anchor = hs; d.has_end_anchor = he;
  if (els.empty()) { b.state(true); return; }
  if (els[0].branches.empty() && !els[0].cc.empty()) b.matcher(els[0].cc);  // not scan eligible (:643-644)
  int cur = 0;
  const int n = (int)els.size();
  for (int i = 0; i < n; ++i) {
    const SeqEl &el = els[i];
    const bool last = i == n - 1;
    bool rest_optional = true;
    for (int k = i + 1; k < n; ++k) if (els[k].mn > 0) { rest_optional = false; break; }
    if (!el.branches.empty()) {  // :660-697
      if (i == 0) { b.state(false); cur = 0; }
      int end_state = b.state(last || rest_optional);
      for (const std::string &br : el.branches) {
        int prev = cur;
        for (size_t ci = 0; ci < br.size(); ++ci) {
          if (ci + 1 == br.size()) b.edge(prev, (uint8_t)br[ci], end_state);
          else { int mid = b.state(false); b.edge(prev, (uint8_t)br[ci], mid); prev = mid; }
        }
      }
      cur = end_state;
      continue;
    }
    if (el.mn == 0) {  // :699-744
      if (i == 0) { b.state(rest_optional); cur = 0; }
      int ms = b.state(last || rest_optional);
      b.cls(cur, ms, el.cc, el.positive);
      if (el.mx == -1) b.cls(ms, ms, el.cc, el.positive);
      cur = (i == 0) ? 0 : ms;
    } else if (el.mn == 1) {  // :746-807
      if (i == 0) { b.state(false); cur = 0; }
      int ms = b.state(last || rest_optional);
      b.cls(cur, ms, el.cc, el.positive);
      if (i == 1 && cur == 0 && els[0].mn == 0) b.cls(1, ms, el.cc, el.positive);
      if (el.mx == -1) b.cls(ms, ms, el.cc, el.positive);
      else if (el.mx > 1) for (int k = 2; k <= el.mx; ++k) { int ex = b.state(last); b.cls(ms + k - 2, ex, el.cc, el.positive); }
      cur = ms;
    } else {  // :809-869
      if (i == 0) { b.state(false); cur = 0; }
      for (int k = 0; k < el.mn; ++k) {
        int si = b.state(k >= el.mn - 1 && last);
        b.cls(k > 0 ? si - 1 : cur, si, el.cc, el.positive);
        cur = si;
      }
      if (el.mx == -1) b.cls(cur, cur, el.cc, el.positive);
      else if (el.mx > el.mn) for (int k = el.mn + 1; k <= el.mx; ++k) { int oi = b.state(last); b.cls(cur, oi, el.cc, el.positive); cur = oi; }
    }
  }
}

// compile_quantified_group :982-1074 (epsilon=true) / compile_simple_quantifier :1147-1243
void build_looping_literal(Builder &b, const std::string &text, int mn, int mx, int acc, bool epsilon_on_zero) {
  RawDfa &d = b.d;
  auto chain = [&](int last_target_or_neg, bool loop_variant) {
    (void)loop_variant;
    int cur = 0;
    for (size_t i = 0; i < text.size(); ++i) {
      int c = (uint8_t)text[i];
      if (i + 1 == text.size()) b.edge(cur, c, last_target_or_neg);
      else { int ns = b.state(); b.edge(cur, c, ns); cur = ns; }
    }
  };
  if (mn == 0 && mx == 1) {
    if (epsilon_on_zero) b.edge(0, 0, acc);  // "epsilon" encoded as a NUL-byte edge (:989)
    else d.accepting[0] = 1;
    chain(acc, false);
  } else if (mn == 0 && mx == -1) {
    d.accepting[0] = 1;
    chain(0, false);
  } else if (mn == 1 && mx == -1) {
    int cur = 0;
    for (size_t i = 0; i < text.size(); ++i) {
      int c = (uint8_t)text[i];
      if (i + 1 == text.size()) {
        int loop = b.state(true);
        b.edge(cur, c, loop);
        b.edge(loop, (uint8_t)text[0], text.size() > 1 ? 1 : loop);
      } else { int ns = b.state(); b.edge(cur, c, ns); cur = ns; }
    }
  } else {
    throw UnsupportedError("Unsupported quantifier range for group");
  }
}

void compile_dfa(const Ast &a, RawDfa &d) {  // compile_dfa_pattern, dfa.mojo:2385-2496
  Builder b(d);
  Shapes sh(a);
  bool hs, he;
  pattern_anchors(a, hs, he);
  const Node &root = a.at(a.root);
  if (is_literal_pattern(a)) {
    build_literal(b, root.kids.empty() ? "" : literal_chars(a, root.kids[0]), hs, he);
  } else if (sh.pure_anchor(a.root)) {
    build_literal(b, "", hs, he);
  } else if (sh.simple_class()) {  // _extract_character_class_info :2532-2602
    int c = root.kids[0];
    int cn = (a.at(c).kind == K_GROUP && a.at(c).kids.size() == 1) ? a.at(c).kids[0] : c;
    const Node &n = a.at(cn);
    std::string cc = n.kind == K_DIGIT ? kDigits : n.kind == K_WORD ? kWord : (a.has_value(cn) ? a.value(cn) : "");
    build_single_class(b, expand_range(cc), n.mn, n.mx, n.positive);
    d.has_start_anchor = hs; d.has_end_anchor = he;
  } else if (sh.multi_class_seq()) {
    build_multi_class(b, sh.multi_class_info(), hs, he);
  } else if (sh.sequential_class()) {  // _extract_sequential_pattern_info :2666-2708
    std::vector<SeqEl> els;
    for (int c : a.at(root.kids[0]).kids) {
      const Node &e = a.at(c);
      std::string cc = e.kind == K_DIGIT ? kDigits : e.kind == K_WORD ? kWord : expand_range(a.value(c));
      els.push_back({cc, e.mn, e.mx, e.positive, {}});
    }
    build_sequential(b, els, hs, he);
  } else if (sh.mixed_sequential()) {
    build_multi_class(b, sh.multi_class_info(), hs, he);
  } else if (sh.alternation()) {  // compile_alternation :873-928
    b.reset();
    b.state(false);
    int orn = sh.find_or(a.root);
    if (orn < 0) throw UnsupportedError("No OR node found in alternation pattern");
    int acc = b.state(true);
    std::vector<std::string> branches;
    sh.all_alt_branches(orn, branches);
    for (const std::string &t : branches) {
      int cur = 0;
      for (size_t j = 0; j < t.size(); ++j) {
        if (j + 1 == t.size()) b.edge(cur, (uint8_t)t[j], acc);
        else cur = b.find_or_create(cur, (uint8_t)t[j]);
      }
    }
    d.has_start_anchor = hs; d.has_end_anchor = he;
  } else if (sh.quantified_group()) {  // :930-980
    b.reset();
    b.state(false);
    int in = sh.quantified_inner_group();
    std::string text;
    for (int c : a.at(in).kids) if (a.at(c).kind == K_ELEMENT) text += a.value(c);
    if (text.empty()) throw UnsupportedError("Empty quantified group");
    int acc = b.state(true);
    build_looping_literal(b, text, a.at(in).mn, a.at(in).mx, acc, true);
    d.has_start_anchor = hs; d.has_end_anchor = he;
  } else if (sh.simple_quantifier()) {  // :1076-1145 (last quantifier applies to the whole string)
    b.reset();
    b.state(false);
    int qmn = 1, qmx = 1;
    std::string text;
    for (int c : a.at(root.kids[0]).kids) {
      const Node &e = a.at(c);
      if (e.mn == 0 && e.mx == -1) { qmn = 0; qmx = -1; }
      else if (e.mn == 1 && e.mx == -1) { qmn = 1; qmx = -1; }
      else if (e.mn == 0 && e.mx == 1) { qmn = 0; qmx = 1; }
      text += a.value(c);
    }
    if (text.empty()) throw UnsupportedError("Empty quantifier pattern");
    int acc = b.state(true);
    build_looping_literal(b, text, qmn, qmx, acc, false);
    d.has_start_anchor = hs; d.has_end_anchor = he;
  } else if (sh.wildcard_quantifier()) {  // :1270-1361
    b.reset();
    b.state(false);
    const Node &w = a.at(a.at(root.kids[0]).kids[0]);
    int acc = b.state(true);
    auto any = [&](int from, int to) { for (int i = 0; i < 256; ++i) if (i != '\n') b.edge(from, i, to); };
    if (w.mn == 0 && w.mx == 1) { d.accepting[0] = 1; any(0, acc); }
    else if (w.mn == 0 && w.mx == -1) { d.accepting[0] = 1; any(0, 0); }
    else if (w.mn == 1 && w.mx == -1) { int loop = b.state(true); any(0, loop); any(loop, loop); }
    else any(0, acc);
    d.has_start_anchor = hs; d.has_end_anchor = he;
  } else if (sh.double_group_or() >= 0 && [&] {  // _is_common_prefix_alternation_pattern :3418-3459
               std::vector<std::string> br;
               Classifier cl(a);
               return cl.strict_branches(sh.double_group_or(), br) && br.size() >= 2 && Classifier::common_prefix(br).size() >= 2;
             }()) {  // compile_common_prefix_alternation :1363-1463
    b.reset();
    b.state(false);
    std::vector<std::string> br;
    sh.lenient_branches(sh.double_group_or(), br);
    if (!br.empty()) {
      std::string prefix = Classifier::common_prefix(br);
      int cur = 0;
      for (unsigned char c : prefix) cur = b.find_or_create(cur, c);
      for (const std::string &s : br) {
        if (s.size() == prefix.size()) { d.accepting[cur] = 1; continue; }
        int sc = cur;
        for (size_t j = prefix.size(); j < s.size(); ++j) {
          if (j + 1 == s.size()) { int t = b.find_or_create(sc, (uint8_t)s[j]); d.accepting[t] = 1; }
          else sc = b.find_or_create(sc, (uint8_t)s[j]);
        }
      }
    }
    d.has_start_anchor = hs; d.has_end_anchor = he;
  } else if (sh.double_group_or() >= 0 && [&] {  // _is_quantified_alternation_group :3525-3562
               const Node &qg = a.at(sh.quantified_inner_group());
               if (qg.mn == 1 && qg.mx == 1) return false;
               std::vector<std::string> br;
               return Classifier(a).strict_branches(sh.double_group_or(), br);
             }()) {  // compile_quantified_alternation_group :1507-1686
    b.reset();
    b.state(false);
    const Node &qg = a.at(sh.quantified_inner_group());
    std::vector<std::string> br;
    sh.lenient_branches(sh.double_group_or(), br);
    auto paths = [&](int origin, int target) {
      for (const std::string &s : br) {
        int cur = origin;
        for (size_t j = 0; j < s.size(); ++j) {
          if (j + 1 == s.size()) b.edge(cur, (uint8_t)s[j], target);
          else cur = b.find_or_create(cur, (uint8_t)s[j]);
        }
      }
    };
    if (qg.mn == 0 && qg.mx == 1) { d.accepting[0] = 1; int acc = b.state(true); paths(0, acc); }
    else if (qg.mn == 0 && qg.mx == -1) { d.accepting[0] = 1; paths(0, 0); }
    else if (qg.mn == 1 && qg.mx == -1) { int loop = b.state(true); paths(0, loop); paths(loop, loop); }
    else throw UnsupportedError("Unsupported quantifier type for quantified alternation group");
    d.has_start_anchor = hs; d.has_end_anchor = he;
  } else {
    throw UnsupportedError("Pattern too complex for current DFA implementation");
  }
  d.start_state = 0;
}

// ---- the NFA side: PikeVM program -> LazyDFA state sets, built eagerly -------------------------------------------
// Patterns without a DFAEngine go to NFAMatcher (matcher.mojo:273-431).  Its match_first / match_next / match_all all
// run on the LazyDFA (pikevm.mojo:682-987: leftmost-longest over the Thompson state sets, cached on demand) exactly when
// the PikeVM program fits (pikevm.mojo:364-366), has no `$` (matcher.mojo:341-350) and the backtracking engine has no
// fast path of its own (literal, leading / trailing `.*`: matcher.mojo:362-385, nfa.mojo:96-139).  For those patterns
// the LazyDFA's reachable states are built here by subset construction and handed on as an ordinary table: an attempt
// at a position is "walk, remember the last accepting position" (pikevm.mojo:806-850) -- the same loop the kernels run
// for a DFAEngine table -- and the callers' loops are the every-position ones (pikevm.mojo:757-803).  `^` is refused:
// the LazyDFA builds its start state at position 0 (pikevm.mojo:714), so the anchor would hold at every start position.
enum PikeOp { P_BYTE, P_RANGE, P_CLASS, P_ANY, P_SPLIT, P_JUMP, P_MATCH, P_START, P_END };
struct PikeInst { int op, a0, a1; };
struct PikeProgram {
  std::vector<PikeInst> ins;
  std::vector<std::array<uint8_t, 256>> classes;
  int emit(int op, int a0 = 0, int a1 = 0) { ins.push_back({op, a0, a1}); return (int)ins.size() - 1; }
  int add_class(const std::array<uint8_t, 256> &t) {  // pikevm.mojo:102-111
    for (size_t i = 0; i < classes.size(); ++i) if (classes[i] == t) return (int)i;
    classes.push_back(t);
    return (int)classes.size() - 1;
  }
};
struct PikeCompiler {
  const Ast &a;
  PikeProgram &p;
  void node(int id) {  // pikevm.mojo:142-166
    const Node &n = a.at(id);
    switch (n.kind) {
      case K_GROUP: for (int c : n.kids) quantified(c); break;  // :169-173
      case K_OR: alt(id); break;
      case K_ELEMENT: if (a.has_value(id)) p.emit(P_BYTE, (uint8_t)a.value(id)[0]); break;  // :260-266
      case K_DIGIT: case K_WORD: case K_SPACE: case K_RANGE: char_class(id); break;
      case K_WILDCARD: p.emit(P_ANY); break;
      case K_START: p.emit(P_START); break;
      case K_END: p.emit(P_END); break;
      default: break;
    }
  }
  void quantified(int id) {  // :176-233
    const int mn = a.at(id).mn, mx = a.at(id).mx;
    if (mn == 1 && mx == 1) { node(id); return; }
    if (mn == 0 && mx == 1) {
      const int sp = p.emit(P_SPLIT);
      const int body = (int)p.ins.size();
      node(id);
      p.ins[sp].a0 = body; p.ins[sp].a1 = (int)p.ins.size();
      return;
    }
    if (mn == mx && mn > 1) { for (int i = 0; i < mn; ++i) node(id); return; }
    if (mx > 0) {
      for (int i = 0; i < mn; ++i) node(id);
      std::vector<int> splits;
      for (int i = 0; i < mx - mn; ++i) { splits.push_back(p.emit(P_SPLIT)); node(id); }
      const int after = (int)p.ins.size();
      for (int s2 : splits) { p.ins[s2].a0 = s2 + 1; p.ins[s2].a1 = after; }
      return;
    }
    if (mx == -1) {
      for (int i = 0; i < mn; ++i) node(id);
      const int sp = p.emit(P_SPLIT);
      const int body = (int)p.ins.size();
      node(id);
      p.emit(P_JUMP, sp);
      p.ins[sp].a0 = body; p.ins[sp].a1 = (int)p.ins.size();
    }
  }
  void alt(int id) {  // :236-257 (children without their quantifiers; the first two only)
    const Node &n = a.at(id);
    if (n.kids.empty()) return;
    if (n.kids.size() == 1) { node(n.kids[0]); return; }
    const int sp = p.emit(P_SPLIT);
    const int left = (int)p.ins.size();
    node(n.kids[0]);
    const int jp = p.emit(P_JUMP);
    const int right = (int)p.ins.size();
    node(n.kids[1]);
    p.ins[sp].a0 = left; p.ins[sp].a1 = right;
    p.ins[jp].a0 = (int)p.ins.size();
  }
  void char_class(int id) {  // :274-333
    const Node &n = a.at(id);
    std::array<uint8_t, 256> t{};
    auto add = [&](const std::string &m) { for (unsigned char c : m) t[c] = 1; };
    if (n.kind == K_DIGIT) add(kDigits);
    else if (n.kind == K_WORD) add(kWord);
    else if (n.kind == K_SPACE) add(" \t\n\r\f");
    else if (a.has_value(id)) {
      std::string raw = a.value(id);
      std::string in = (raw.size() >= 2 && raw.front() == '[' && raw.back() == ']') ? raw.substr(1, raw.size() - 2) : raw;
      size_t j = 0;
      while (j < in.size()) {
        if (j + 1 < in.size() && in[j] == '\\') {
          const char nx = in[j + 1];
          if (nx == 's') add(" \t\n\r\f");
          else if (nx == 'd') add(kDigits);
          else if (nx == 'w') add(kWord);
          else t[(uint8_t)nx] = 1;
          j += 2;
        } else if (j + 2 < in.size() && in[j + 1] == '-') {
          for (int c = (uint8_t)in[j]; c <= (int)(uint8_t)in[j + 2]; ++c) t[c] = 1;
          j += 3;
        } else {
          t[(uint8_t)in[j]] = 1;  // (a leading '^' of a negated class is a member too: the value is the raw slice)
          ++j;
        }
      }
    }
    if (!n.positive) for (int c = 0; c < 256; ++c) t[c] = 1 - t[c];
    p.emit(P_CLASS, p.add_class(t));
  }
};

// epsilon closure (PikeVMEngine._add_state, pikevm.mojo:604-653) of a set of entry pcs; anchors never hold (programs
// that have them are refused before this is called).  Result: sorted leaf pcs.
std::vector<int> pike_closure(const PikeProgram &p, const std::vector<int> &entry) {
  std::vector<uint8_t> seen(p.ins.size(), 0);
  std::vector<int> stack(entry), leaves;
  while (!stack.empty()) {
    const int pc = stack.back();
    stack.pop_back();
    if (pc >= (int)p.ins.size() || seen[pc]) continue;
    const PikeInst &in = p.ins[pc];
    if (in.op == P_START || in.op == P_END) continue;
    seen[pc] = 1;
    if (in.op == P_SPLIT) { stack.push_back(in.a0); stack.push_back(in.a1); }
    else if (in.op == P_JUMP) stack.push_back(in.a0);
    else leaves.push_back(pc);
  }
  std::sort(leaves.begin(), leaves.end());
  return leaves;
}
bool pike_consumes(const PikeProgram &p, int pc, int ch) {
  const PikeInst &in = p.ins[pc];
  switch (in.op) {
    case P_BYTE: return ch == in.a0;
    case P_CLASS: return p.classes[in.a0][ch] != 0;
    case P_ANY: return ch != 10;
    case P_RANGE: return ch >= in.a0 && ch <= in.a1;
    default: return false;
  }
}

// fills `d` with the LazyDFA of the pattern, or throws UnsupportedError saying which other engine the reference uses
void compile_lazy_dfa(const Ast &ast, const std::string &pattern, const std::string &why, RawDfa &d) {
  PikeProgram prog;
  PikeCompiler pc{ast, prog};
  const Node &root = ast.at(ast.root);
  if (root.kind == K_RE && !root.kids.empty()) pc.node(root.kids[0]);  // pikevm.mojo:124-139
  prog.emit(P_MATCH);
  if (prog.ins.size() > 512) throw UnsupportedError(why + "; PikeVM program above 512 instructions: backtracking NFA");
  for (const PikeInst &in : prog.ins) {
    if (in.op == P_END) throw UnsupportedError(why + "; `$` in an NFA-routed pattern: served by OnePass / the backtracking NFA");
    if (in.op == P_START)
      throw UnsupportedError(why + "; `^` in an NFA-routed pattern: the LazyDFA builds its start state at position 0 (pikevm.mojo:714)");
  }
  Lit best;
  if (best_literal(ast, best) &&  // NFAEngine.__init__, nfa.mojo:96-126
      ((best.prefix && best.required && best.text.size() >= 1) || (best.required && best.text.size() >= 3)))
    throw UnsupportedError(why + "; the backtracking NFA has a literal fast path for this pattern: it serves match_next / match_all");
  auto ends_with = [&](const std::string &x) { return pattern.size() >= x.size() && pattern.compare(pattern.size() - x.size(), x.size(), x) == 0; };
  const bool ends = ends_with(".*") && !ends_with("\\.*");  // nfa.mojo:129-139
  const bool starts = pattern.compare(0, 2, ".*") == 0 && !(pattern.size() > 2 && (pattern[2] == '?' || pattern[2] == '*' || pattern[2] == '+'));
  if (ends || starts) throw UnsupportedError(why + "; leading / trailing `.*`: the backtracking NFA's fast path serves it");
  // subset construction over the sets the LazyDFA would cache
  std::map<std::vector<int>, int> index;
  std::vector<std::vector<int>> order;
  order.push_back(pike_closure(prog, {0}));
  index[order[0]] = 0;
  d = RawDfa();
  for (size_t i = 0; i < order.size(); ++i) {
    const std::vector<int> cur = order[i];
    std::array<int32_t, 256> row;
    row.fill(-1);
    std::map<std::vector<int>, int> move_state;  // move set -> DFA state
    for (int ch = 0; ch < 256; ++ch) {
      std::vector<int> mv;
      for (int q : cur) if (pike_consumes(prog, q, ch)) mv.push_back(q + 1);
      if (mv.empty()) continue;
      auto it = move_state.find(mv);
      if (it == move_state.end()) {
        const std::vector<int> nxt = pike_closure(prog, mv);
        int id = -1;
        if (!nxt.empty()) {  // an empty set is LAZY_DFA_DEAD (pikevm.mojo:923-924)
          auto f = index.find(nxt);
          if (f == index.end()) {
            if (order.size() >= 254) throw UnsupportedError(why + "; more than 254 LazyDFA states");
            id = (int)order.size();
            index[nxt] = id;
            order.push_back(nxt);
          } else {
            id = f->second;
          }
        }
        it = move_state.emplace(mv, id).first;
      }
      row[ch] = it->second;
    }
    d.trans.push_back(row);
  }
  for (const std::vector<int> &st : order) {
    bool acc = false;
    for (int q : st) acc = acc || prog.ins[q].op == P_MATCH;
    d.accepting.push_back(acc ? 1 : 0);
  }
  d.start_state = 0;
}

int rare_required_byte(const Ast &a, const std::array<uint8_t, 256> &lookup) {  // matcher.mojo:122-165
  int id = a.root;
  if (a.at(id).kids.size() == 1) id = a.at(id).kids[0];
  if (a.at(id).kind != K_GROUP) return -1;
  for (int c : a.at(id).kids) {
    const Node &e = a.at(c);
    if (e.kind != K_ELEMENT || e.mn < 1 || !a.has_value(c)) continue;
    std::string v = a.value(c);
    if (v.size() != 1) continue;
    if (!lookup[(uint8_t)v[0]]) return (uint8_t)v[0];
  }
  return -1;
}

bool fixed_width_groups(const std::string &p, std::vector<int> &segs) {  // matcher.mojo:1486-1586
  const int n = (int)p.size();
  int lit = 0;
  for (int i = 0; i < n;) {
    if (p[i] == '(') {
      if (lit > 0) { segs.push_back(-lit); lit = 0; }
      if (i + 1 < n && p[i + 1] == '?') return false;
      ++i;
      if (i + 1 >= n || p[i] != '\\' || p[i + 1] != 'd') return false;
      i += 2;
      if (i < n && p[i] == '{') {
        ++i;
        int ns = i, w = 0;
        while (i < n && p[i] >= '0' && p[i] <= '9') { w = w * 10 + (p[i] - '0'); ++i; }
        if (i == ns || i >= n || p[i] != '}') return false;
        ++i;
        segs.push_back(w);
      } else if (i < n && p[i] == ')') {
        segs.push_back(1);
      } else {
        return false;
      }
      if (i >= n || p[i] != ')') return false;
      ++i;
    } else if (p[i] == '|' || p[i] == '[') {
      return false;
    } else if (p[i] == '\\' && i + 1 < n) { ++lit; i += 2; }
    else { ++lit; ++i; }
  }
  for (int s : segs) if (s > 0) return true;
  return false;
}

void precompute_fixed_sub(const std::string &pattern, RegexMeta &m) {  // matcher.mojo:1002-1035
  std::vector<int> segs;
  if (!fixed_width_groups(pattern, segs)) return;
  int ng = 0, tw = 0;
  bool has_lit = false;
  int offs[10] = {0}, wid[10] = {0};
  for (int s : segs) {
    if (s > 0) {
      if (++ng > 9) return;
      offs[ng] = tw; wid[ng] = s; tw += s;
    } else { has_lit = true; tw += -s; }
  }
  if (ng == 0) return;
  std::memcpy(m.fixed_offsets, offs, sizeof offs);
  std::memcpy(m.fixed_widths, wid, sizeof wid);
  m.fixed_num_groups = ng; m.fixed_total_width = tw; m.fixed_concat = !has_lit;
}

}  // namespace

int compile_frontend(const std::string &pattern, RawDfa &dfa, RegexMeta &meta, std::string &err) {
  try {
    dfa = RawDfa();
    meta = RegexMeta();
    if (pattern == ".*") {  // matcher.mojo:573-591
      meta.is_wildcard_any = true;
      precompute_fixed_sub(pattern, meta);
      return MRX_OK;
    }
    Ast ast = parse(pattern);
    Classifier cl(ast);
    const int complexity = cl.node(ast.root, 0);
    const bool pure_dfa = complexity == C_SIMPLE && cl.simd_nodes(ast.root) <= 1;  // optimizer.mojo:174-201
    if (!skip_prefilter(pattern) && !pure_dfa) {  // matcher.mojo:604-654
      meta.has_anchors_info = ast_has_anchor(ast, ast.root);
      Lit best;
      if (best_literal(ast, best) && !best.text.empty()) {
        meta.exact_literal = best.text;
        bool ops = pattern.find_first_of("*+?.|([{") != std::string::npos;
        const Node &r = ast.at(ast.root);
        bool lit_prefix = !r.kids.empty() && literal_prefix_node(ast, r.kids[0]);
        meta.is_exact_literal = best.required && lit_prefix && !meta.has_anchors_info && !ops;
        if (pattern.find('|') == std::string::npos && best.text.size() >= 2) meta.prefilter = best.text;
      }
    }
    bool nfa_routed = false;
    if (complexity != C_SIMPLE) {  // matcher.mojo:665-675: no DFAEngine -> NFAMatcher; served when that means its LazyDFA
      compile_lazy_dfa(ast, pattern, std::string("pattern classified ") + (complexity == C_MEDIUM ? "MEDIUM" : "COMPLEX") + " -> NFA engines", dfa);
      nfa_routed = true;
    } else {
      try {
        compile_dfa(ast, dfa);
      } catch (const UnsupportedError &e) {  // DFA compilation raised: NFA only (matcher.mojo:668-672)
        compile_lazy_dfa(ast, pattern, e.what(), dfa);
        nfa_routed = true;
      }
    }
    if (!nfa_routed && !meta.has_anchors_info && dfa.has_matcher) meta.required_byte = rare_required_byte(ast, dfa.lookup);  // :684-693
    precompute_fixed_sub(pattern, meta);
    return MRX_OK;
  } catch (const ParseError &e) {
    err = e.what();
    return MRX_E_PARSE;
  } catch (const UnsupportedError &e) {
    err = e.what();
    return MRX_E_UNSUPPORTED;
  }
}

}  // namespace mrx
