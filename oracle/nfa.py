"""ORACLE (test infrastructure, NOT product code) -- the reference's NFA side, as far as the GPU path serves it.

Patterns the reference does not hand to its DFAEngine (classification MEDIUM / COMPLEX, or `compile_dfa_pattern`
raising, matcher.mojo:665-675) go to `NFAMatcher` (matcher.mojo:273-431), which owns three engines:

  * `LazyDFA` over the PikeVM program (pikevm.mojo:682-987): leftmost-longest over the Thompson state sets,
    `match_next` = the first start position whose run reaches MATCH, `match_all` restarts at the match end
    (one past the start after an empty match);
  * `OnePassNFA` for `$`-anchored programs;
  * the backtracking `NFAEngine` with its literal / `.*` fast paths.

`match_next` / `match_all` use the LazyDFA exactly when it exists and the backtracking engine has no fast path
(`_use_lazy_dfa_for_search`, matcher.mojo:383-385, 362-371); `match_first` uses it whenever it exists and the
program has no `$` (matcher.mojo:341-350).  This module restates the PikeVM compiler (pikevm.mojo:124-333), that
routing predicate, the state-set semantics (`run`, a direct simulation used by the tests) and the EAGER subset
construction whose table goes through the same scan loops as a DFAEngine table (the LazyDFA builds the same states
on demand).  Everything else -- `$` programs (OnePass / backtracking), patterns with a literal or `.*` fast path,
and `^` (the LazyDFA builds its start state at position 0, pikevm.mojo:714, so the anchor would hold at every start
position; not reproduced) -- stays `Unsupported`: the caller keeps its CPU engine.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, FrozenSet, List, Optional, Tuple

OP_BYTE, OP_RANGE, OP_CLASS, OP_ANY, OP_SPLIT, OP_JUMP, OP_MATCH, OP_START_ANCHOR, OP_END_ANCHOR = range(9)
MAX_STATES = 512  # pikevm.mojo:342
MIN_PREFIX_LITERAL_LENGTH = 1    # nfa.mojo:61
MIN_REQUIRED_LITERAL_LENGTH = 3  # nfa.mojo:63


@dataclass
class Program:
    """pikevm.mojo:81-117."""
    instrs: List[List[int]] = field(default_factory=list)  # [opcode, arg0, arg1]
    class_tables: List[bytes] = field(default_factory=list)  # 256 bytes of 0 / 1

    def emit(self, op: int, a0: int = 0, a1: int = 0) -> int:
        self.instrs.append([op, a0, a1])
        return len(self.instrs) - 1

    def add_class_table(self, table: bytes) -> int:  # :102-111 (identical tables are shared)
        for i, t in enumerate(self.class_tables):
            if t == table:
                return i
        self.class_tables.append(table)
        return len(self.class_tables) - 1

    def patch(self, idx: int, a0: int = -1, a1: int = -1) -> None:  # :113-118
        if a0 >= 0:
            self.instrs[idx][1] = a0
        if a1 >= 0:
            self.instrs[idx][2] = a1


def compile_program(ast, F) -> Program:
    """pikevm.mojo:124-139.  `F` = the oracle's frontend module (node kinds, class constants)."""
    p = Program()
    if ast.type == F.RE and ast.get_children_len() > 0:
        _node(ast.get_child(0), p, F)
    p.emit(OP_MATCH)
    return p


def _node(node, p: Program, F) -> None:  # :142-166
    t = node.type
    if t == F.GROUP:
        for child in node.children:  # :169-173
            _quantified(child, p, F)
    elif t == F.OR:
        _or(node, p, F)
    elif t == F.ELEMENT:
        v = node.get_value()  # :260-266
        if v:
            p.emit(OP_BYTE, v[0])
    elif t in (F.DIGIT, F.WORD, F.SPACE, F.RANGE):
        _char_class(node, p, F)
    elif t == F.WILDCARD:
        p.emit(OP_ANY)
    elif t == F.START:
        p.emit(OP_START_ANCHOR)
    elif t == F.END:
        p.emit(OP_END_ANCHOR)


def _quantified(node, p: Program, F) -> None:
    """pikevm.mojo:176-233."""
    mn, mx = node.min, node.max
    if mn == 1 and mx == 1:
        _node(node, p, F)
        return
    if mn == 0 and mx == 1:
        sp = p.emit(OP_SPLIT, 0, 0)
        body = len(p.instrs)
        _node(node, p, F)
        p.patch(sp, a0=body, a1=len(p.instrs))
        return
    if mn == mx and mn > 1:
        for _ in range(mn):
            _node(node, p, F)
        return
    if mx > 0:
        for _ in range(mn):
            _node(node, p, F)
        splits = []
        for _ in range(mx - mn):
            splits.append(p.emit(OP_SPLIT, 0, 0))
            _node(node, p, F)
        after = len(p.instrs)
        for s in splits:
            p.patch(s, a0=s + 1, a1=after)
        return
    if mx == -1:
        for _ in range(mn):
            _node(node, p, F)
        sp = p.emit(OP_SPLIT, 0, 0)
        body = len(p.instrs)
        _node(node, p, F)
        p.emit(OP_JUMP, sp)
        p.patch(sp, a0=body, a1=len(p.instrs))


def _or(node, p: Program, F) -> None:
    """pikevm.mojo:236-257 (the children are compiled WITHOUT their quantifiers, and only the first two count)."""
    k = node.get_children_len()
    if k == 0:
        return
    if k == 1:
        _node(node.get_child(0), p, F)
        return
    sp = p.emit(OP_SPLIT, 0, 0)
    left = len(p.instrs)
    _node(node.get_child(0), p, F)
    jp = p.emit(OP_JUMP, 0)
    right = len(p.instrs)
    _node(node.get_child(1), p, F)
    p.patch(sp, a0=left, a1=right)
    p.patch(jp, a0=len(p.instrs))


def _char_class(node, p: Program, F) -> None:
    """pikevm.mojo:274-333."""
    table = bytearray(256)
    if node.type == F.DIGIT:
        members = F.DIGITS
    elif node.type == F.WORD:
        members = F.WORD_CHARS
    elif node.type == F.SPACE:
        members = b" \t\n\r\x0c"  # dfa.mojo:88-89
    else:
        members = b""
    for b in members:
        table[b] = 1
    if node.type == F.RANGE and node.get_value():
        raw = node.get_value()
        inner = raw[1:-1] if (raw.startswith(b"[") and raw.endswith(b"]")) else raw
        j, n = 0, len(inner)
        while j < n:
            if j + 1 < n and inner[j] == ord("\\"):
                nx = inner[j + 1]
                if nx == ord("s"):
                    for b in b" \t\n\r\x0c":
                        table[b] = 1
                elif nx == ord("d"):
                    for b in range(ord("0"), ord("9") + 1):
                        table[b] = 1
                elif nx == ord("w"):
                    for b in F.WORD_CHARS:
                        table[b] = 1
                else:
                    table[nx] = 1
                j += 2
            elif j + 2 < n and inner[j + 1] == ord("-"):
                for b in range(inner[j], inner[j + 2] + 1):
                    if b < 256:
                        table[b] = 1
                j += 3
            else:
                table[inner[j]] = 1  # (a leading '^' of a negated class is a member too: the value is the raw slice)
                j += 1
    if not node.positive_logic:
        for b in range(256):
            table[b] = 1 - table[b]
    p.emit(OP_CLASS, p.add_class_table(bytes(table)))


# --------------------------------------------------------------------------------------------------------------
def closure(p: Program, pcs, pos: int, text_len: int) -> FrozenSet[int]:
    """PikeVMEngine._add_state (pikevm.mojo:604-653) for a set of entry pcs: the leaf instructions reached."""
    leaves, seen, stack = set(), set(), list(pcs)
    n = len(p.instrs)
    while stack:
        pc = stack.pop()
        if pc >= n or pc in seen:
            continue
        op, a0, a1 = p.instrs[pc]
        if op == OP_SPLIT:
            seen.add(pc)
            stack += [a0, a1]
        elif op == OP_JUMP:
            seen.add(pc)
            stack.append(a0)
        elif op == OP_START_ANCHOR:
            if pos == 0:
                seen.add(pc)
                stack.append(pc + 1)
        elif op == OP_END_ANCHOR:
            if pos == text_len:
                seen.add(pc)
                stack.append(pc + 1)
        else:
            seen.add(pc)
            leaves.add(pc)
    return frozenset(leaves)


def _consumes(p: Program, pc: int, ch: int) -> bool:
    op, a0, a1 = p.instrs[pc]
    if op == OP_BYTE:
        return ch == a0
    if op == OP_CLASS:
        return p.class_tables[a0][ch] != 0
    if op == OP_ANY:
        return ch != 10
    if op == OP_RANGE:
        return a0 <= ch <= a1
    return False


def run(p: Program, text: bytes, start: int) -> Optional[Tuple[int, int]]:
    """PikeVMEngine._run (pikevm.mojo:497-602): the longest match that begins at `start`, by direct simulation of
    the state sets (the LazyDFA caches exactly these sets).  Programs without anchors only."""
    n = len(text)
    cur = closure(p, [0], 1, -1)  # anchors never hold (callers refuse programs that have them)
    match_end = -1
    pos = start
    while pos <= n and cur:
        if any(p.instrs[pc][0] == OP_MATCH for pc in cur):
            match_end = pos
        if pos == n:
            break
        ch = text[pos]
        cur = closure(p, [pc + 1 for pc in cur if _consumes(p, pc, ch)], 1, -1)
        pos += 1
    return (start, match_end) if match_end >= 0 else None


def has_literal_optimization(ast, F) -> bool:
    """NFAEngine.__init__ (nfa.mojo:96-126)."""
    best = F.best_literal(ast)
    if best is None:
        return False
    if best.is_prefix and best.is_required and len(best.literal) >= MIN_PREFIX_LITERAL_LENGTH:
        return True
    return bool(best.is_required and len(best.literal) >= MIN_REQUIRED_LITERAL_LENGTH)


def route(pattern: bytes, ast, F) -> Tuple[Optional[Program], str]:
    """-> (program, "") when match_first, match_next and match_all all run on the LazyDFA, else (None, why)."""
    p = compile_program(ast, F)
    if len(p.instrs) > MAX_STATES:  # PikeVMEngine.is_supported, pikevm.mojo:364-366
        return None, "PikeVM program above 512 instructions: backtracking NFA"
    ops = {i[0] for i in p.instrs}
    if OP_END_ANCHOR in ops:  # LazyDFA.has_end_anchor: OnePass or the backtracking engine (matcher.mojo:341-350)
        return None, "`$` in an NFA-routed pattern: served by OnePass / the backtracking NFA"
    if OP_START_ANCHOR in ops:
        return None, "`^` in an NFA-routed pattern: the LazyDFA builds its start state at position 0 (pikevm.mojo:714)"
    if has_literal_optimization(ast, F):  # _nfa_fast_paths_absent, matcher.mojo:362-371
        return None, "the backtracking NFA has a literal fast path for this pattern: it serves match_next / match_all"
    ends = pattern.endswith(b".*") and not pattern.endswith(b"\\.*")  # nfa.mojo:129-131
    starts = pattern.startswith(b".*") and not (len(pattern) > 2 and pattern[2] in b"?*+")  # :132-139
    if ends or starts:
        return None, "leading / trailing `.*`: the backtracking NFA's fast path serves it"
    return p, ""


def subset_construction(p: Program, max_states: int = 254):
    """The LazyDFA's states, built eagerly: (trans[n][256] with -1 = dead, accepting[n], start) or None when there are
    more than `max_states`."""
    start = closure(p, [0], 1, -1)
    index: Dict[FrozenSet[int], int] = {start: 0}
    order = [start]
    trans: List[List[int]] = []
    i = 0
    while i < len(order):
        cur = order[i]
        row = [-1] * 256
        # group the bytes by the set of leaves that consume them
        by_move: Dict[FrozenSet[int], List[int]] = {}
        for ch in range(256):
            mv = frozenset(pc + 1 for pc in cur if _consumes(p, pc, ch))
            if mv:
                by_move.setdefault(mv, []).append(ch)
        for mv, chars in by_move.items():
            nxt = closure(p, mv, 1, -1)
            if not nxt:
                continue  # nxt_count == 0 -> LAZY_DFA_DEAD (pikevm.mojo:923-924)
            if nxt not in index:
                if len(order) >= max_states:
                    return None
                index[nxt] = len(order)
                order.append(nxt)
            for ch in chars:
                row[ch] = index[nxt]
        trans.append(row)
        i += 1
    accepting = [any(p.instrs[pc][0] == OP_MATCH for pc in s) for s in order]
    return trans, accepting, 0
