"""ORACLE (test infrastructure, NOT product code) -- pattern front end.

CPU restatement, in plain Python, of the parts of msaelices/mojo-regex v0.21.0
that turn a pattern string into the hand-built DFA tables and dispatch flags the
matching hot path runs on.  Every function cites the reference file:line it
follows (paths relative to /root/reference/src/regex/).  Compile-side work is a
"small case" (patterns are tens of bytes), hence pure Python; the scan loops that
touch text live in oracle/scan.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (mojo_regex_b200/) has its own,
independently written C++ front end and must never import anything from oracle/.

Parity status: the reference cannot be executed here (no Mojo toolchain), so
this restatement is pinned against the golden vectors of the reference's own
test-suite (tests/golden/reference_vectors.json, transcribed with file:line from
/root/reference/tests/*.mojo by tests/golden/make_reference_vectors.py).  Config
1 and 4 are pinned by reference tests; configs 2, 3 and 5 are pinned only by
adjacent vectors + README -> "parity partially unpinned".
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

# --------------------------------------------------------------------------- #
# aliases.mojo:6-9, dfa.mojo:64-68
DIGITS = b"0123456789"
WORD_CHARS = b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_"
LOWERCASE_LETTERS = b"abcdefghijklmnopqrstuvwxyz"
UPPERCASE_LETTERS = b"ABCDEFGHIJKLMNOPQRSTUVWXYZ"
ALL_LETTERS = LOWERCASE_LETTERS + UPPERCASE_LETTERS
ALPHANUMERIC = LOWERCASE_LETTERS + UPPERCASE_LETTERS + DIGITS
# aliases.mojo:1-3 -- chr(32..126) without '\n' (which is not in that range anyway)
ALL_EXCEPT_NEWLINE = bytes(i for i in range(32, 127) if i != 10)
SPACE_CLASS = b" \t\n\r\x0c"  # dfa.mojo:89, 2901-2902 (no \v in the DFA)

# ast.mojo:25-36
RE, ELEMENT, WILDCARD, SPACE, DIGIT, WORD, RANGE, START, END, OR, NOT, GROUP = range(12)

MAX_LITERAL_QUANT_REPETITIONS = 10  # optimizer.mojo:33
SIMPLE, MEDIUM, COMPLEX = 0, 1, 2  # optimizer.mojo:38-46


class RegexError(Exception):
    """Pattern the reference's parser rejects (parser.mojo:225-237,340,399; lexer.mojo:150)."""


class Unsupported(Exception):
    """Pattern the reference routes to its NFA engines (out of scope: SURVEY 8f)."""


# --------------------------------------------------------------------------- #
# Lexer: lexer.mojo:61-195
T_ELEMENT, T_WILDCARD, T_SPACE, T_DIGIT, T_WORD, T_START, T_END, T_COMMA, \
    T_LPAREN, T_RPAREN, T_LCURLY, T_RCURLY, T_LBRACKET, T_RBRACKET, T_QMARK, \
    T_ASTERISK, T_PLUS, T_VBAR, T_CIRCUMFLEX, T_DASH = range(20)


@dataclass
class Token:
    type: int
    char: int = 0
    start_pos: int = 0


def scan(regex: bytes) -> List[Token]:
    """lexer.mojo:61-195."""
    tokens: List[Token] = []
    i = 0
    escape_found = False
    n = len(regex)
    while i < n:
        ch = regex[i]
        if escape_found:
            if ch == ord("t"):  # lexer.mojo:79-82 (value slice = the backslash byte)
                tokens.append(Token(T_ELEMENT, 9, i - 1))
            elif ch == ord("s"):
                tokens.append(Token(T_SPACE, ch, i - 1))
            elif ch == ord("d"):
                tokens.append(Token(T_DIGIT, ch, i - 1))
            elif ch == ord("w"):
                tokens.append(Token(T_WORD, ch, i - 1))
            else:
                tokens.append(Token(T_ELEMENT, ch, i))
        elif ch == ord("\\"):
            escape_found = True
            i += 1
            continue
        elif ch == ord("."):
            tokens.append(Token(T_WILDCARD, 0, i))
        elif ch == ord("("):
            tokens.append(Token(T_LPAREN, 0, i))
        elif ch == ord(")"):
            tokens.append(Token(T_RPAREN, 0, i))
        elif ch == ord("["):
            tokens.append(Token(T_LBRACKET, 0, i))
        elif ch == ord("-"):
            tokens.append(Token(T_DASH, 0, i))
        elif ch == ord("]"):
            tokens.append(Token(T_RBRACKET, 0, i))
        elif ch == ord("{"):  # lexer.mojo:129-153: braces are lexed eagerly
            tokens.append(Token(T_LCURLY, 0, i))
            i += 1
            while i < n:
                c2 = regex[i]
                if c2 == ord(","):
                    tokens.append(Token(T_COMMA, 0, i))
                elif ord("0") <= c2 <= ord("9"):
                    tokens.append(Token(T_ELEMENT, c2, i))
                elif c2 == ord("}"):
                    tokens.append(Token(T_RCURLY, 0, i))
                    break
                else:
                    raise RegexError("Bad token at index %d." % i)
                i += 1
        elif ch == ord("^"):
            tokens.append(Token(T_START if i == 0 else T_CIRCUMFLEX, 0, i))
        elif ch == ord("$"):
            tokens.append(Token(T_END, 0, i))
        elif ch == ord("?"):
            tokens.append(Token(T_QMARK, 0, i))
        elif ch == ord("*"):
            tokens.append(Token(T_ASTERISK, 0, i))
        elif ch == ord("+"):
            tokens.append(Token(T_PLUS, 0, i))
        elif ch == ord("|"):
            tokens.append(Token(T_VBAR, 0, i))
        elif ch == ord("}"):
            tokens.append(Token(T_RCURLY, 0, i))
        else:
            tokens.append(Token(T_ELEMENT, ch, i))
        escape_found = False
        i += 1
    return tokens


# --------------------------------------------------------------------------- #
# AST: ast.mojo:168-300, 540-558
@dataclass
class Node:
    type: int
    pattern: bytes
    start_idx: int = 0
    end_idx: int = 0
    min: int = 1
    max: int = 1
    positive_logic: bool = True
    capturing_group: bool = False
    group_id: int = -1
    children: List["Node"] = field(default_factory=list)

    def get_value(self) -> Optional[bytes]:
        """ast.mojo:546-558 -- the raw pattern slice; None when empty."""
        if self.start_idx == self.end_idx:
            return None
        return self.pattern[self.start_idx:self.end_idx]

    def get_children_len(self) -> int:
        return len(self.children)

    def get_child(self, i: int) -> "Node":
        return self.children[i]

    def has_children(self) -> bool:
        return len(self.children) > 0


def _check_for_quantifiers(i: int, elem: Node, tokens: List[Token]) -> int:
    """parser.mojo:56-114.  Returns the updated token index."""
    nxt = tokens[i + 1]
    if nxt.type == T_ASTERISK:
        elem.min, elem.max = 0, -1
        i += 1
    elif nxt.type == T_PLUS:
        elem.min, elem.max = 1, -1
        i += 1
    elif nxt.type == T_QMARK:
        elem.min, elem.max = 0, 1
        i += 1
    elif nxt.type == T_LCURLY:
        i += 2
        min_val = max_val = 0
        has_min = has_max = False
        while i < len(tokens) and tokens[i].type == T_ELEMENT:
            d = tokens[i].char
            if ord("0") <= d <= ord("9"):
                min_val = min_val * 10 + (d - ord("0"))
                has_min = True
            else:
                raise RegexError("Invalid digit in quantifier")
            i += 1
        elem.min = min_val if has_min else 0
        if i < len(tokens) and tokens[i].type == T_COMMA:
            i += 1
            while i < len(tokens) and tokens[i].type == T_ELEMENT:
                d = tokens[i].char
                if ord("0") <= d <= ord("9"):
                    max_val = max_val * 10 + (d - ord("0"))
                    has_max = True
                else:
                    raise RegexError("Invalid digit in quantifier")
                i += 1
            elem.max = max_val if has_max else -1
        else:
            elem.max = elem.min
        if i < len(tokens) and tokens[i].type == T_RCURLY:
            i += 1
        i -= 1
    return i


def _group_node(pattern: bytes, children: List[Node], start_idx: int, end_idx: int,
                capturing: bool, group_id: int) -> Node:
    return Node(GROUP, pattern, start_idx, end_idx, 1, 1, True, capturing, group_id, children)


def _parse_token_list(pattern: bytes, tokens: List[Token], counter: List[int]) -> Node:
    """parser.mojo:130-464.  `counter` is the 1-element mutable group counter."""
    if len(tokens) == 0:
        return _group_node(pattern, [], 0, 0, True, 0)

    # Top-level '|' split -> right-nested binary OR (parser.mojo:150-214)
    depth = 0
    for k, tok in enumerate(tokens):
        if tok.type == T_LPAREN:
            depth += 1
        elif tok.type == T_RPAREN:
            depth -= 1
        elif tok.type == T_VBAR and depth == 0:
            left_tokens, right_tokens = tokens[:k], tokens[k + 1:]
            left = (_parse_token_list(pattern, left_tokens, counter) if left_tokens
                    else _group_node(pattern, [], 0, 0, True, 0))
            right = (_parse_token_list(pattern, right_tokens, counter) if right_tokens
                     else _group_node(pattern, [], 0, 0, True, 0))
            return Node(OR, pattern, 0, len(pattern), 1, 1, True, False, -1, [left, right])

    # Validation (parser.mojo:216-237)
    bdepth = pdepth = 0
    for tok in tokens:
        if tok.type == T_LBRACKET:
            bdepth += 1
        elif tok.type == T_RBRACKET:
            bdepth -= 1
            if bdepth < 0:
                raise RegexError("Unescaped closing bracket ']' at position %d" % tok.start_pos)
        elif tok.type == T_LPAREN:
            pdepth += 1
        elif tok.type == T_RPAREN:
            pdepth -= 1
            if pdepth < 0:
                raise RegexError("Unescaped closing parenthesis ')' at position %d" % tok.start_pos)

    elements: List[Node] = []
    i = 0
    n = len(tokens)
    while i < n:
        tok = tokens[i]
        t = tok.type
        if t == T_ELEMENT or t == T_DASH:  # parser.mojo:246-255, 355-364
            elem = Node(ELEMENT, pattern, tok.start_pos, tok.start_pos + 1)
            if i + 1 < n:
                i = _check_for_quantifiers(i, elem, tokens)
            elements.append(elem)
        elif t in (T_WILDCARD, T_SPACE, T_DIGIT, T_WORD):  # parser.mojo:256-298
            kind = {T_WILDCARD: WILDCARD, T_SPACE: SPACE, T_DIGIT: DIGIT, T_WORD: WORD}[t]
            width = 1 if t == T_WILDCARD else 2
            elem = Node(kind, pattern, tok.start_pos, tok.start_pos + width)
            if i + 1 < n:
                i = _check_for_quantifiers(i, elem, tokens)
            elements.append(elem)
        elif t == T_START:
            elements.append(Node(START, pattern, tok.start_pos, tok.start_pos + 1))
        elif t == T_END:
            elements.append(Node(END, pattern, tok.start_pos, tok.start_pos + 1))
        elif t == T_LBRACKET:  # parser.mojo:313-354 -- value is the raw pattern slice
            bracket_start = tok.start_pos
            i += 1
            positive = True
            if i < n and tokens[i].type == T_CIRCUMFLEX:
                positive = False
                i += 1
            while i < n and tokens[i].type != T_RBRACKET:
                if (i + 2 < n and tokens[i + 1].type == T_DASH
                        and tokens[i + 2].type == T_ELEMENT):
                    i += 3
                else:
                    i += 1
            if i >= n:
                raise RegexError("Missing closing ']'.")
            bracket_end = tokens[i].start_pos + 1
            elem = Node(RANGE, pattern, bracket_start, bracket_end, 1, 1, positive)
            if i + 1 < n:
                i = _check_for_quantifiers(i, elem, tokens)
            elements.append(elem)
        elif t == T_LPAREN:  # parser.mojo:365-444
            paren_start = tok.start_pos
            i += 1
            capturing = True
            content_start = paren_start + 1
            if (i + 1 < n and tokens[i].type == T_QMARK
                    and tokens[i + 1].type == T_ELEMENT and tokens[i + 1].char == ord(":")):
                capturing = False
                i += 2
                content_start = paren_start + 3
            group_tokens: List[Token] = []
            count = 1
            while i < n and count > 0:
                if tokens[i].type == T_LPAREN:
                    count += 1
                elif tokens[i].type == T_RPAREN:
                    count -= 1
                    if count == 0:
                        break
                group_tokens.append(tokens[i])
                i += 1
            if count > 0:
                raise RegexError("Missing closing parenthesis ')'.")
            paren_end = tokens[i].start_pos
            gid = -1
            if capturing:
                counter[0] += 1
                gid = counter[0]
            inner = _parse_token_list(pattern, group_tokens, counter)
            if inner.type == GROUP:
                group = inner
                group.capturing_group = capturing
                group.group_id = gid
                group.start_idx = content_start
                group.end_idx = paren_end
            else:
                group = _group_node(pattern, [inner], content_start, paren_end, capturing, gid)
            if i + 1 < n:
                i = _check_for_quantifiers(i, group, tokens)
            elements.append(group)
        # every other token type is silently skipped (no branch in parser.mojo:243-446)
        i += 1

    return _group_node(pattern, elements, 0, len(pattern), True, 0)


def parse(pattern: bytes) -> Node:
    """parser.mojo:467-510 -- RE root wrapping the parsed result."""
    tokens = scan(pattern)
    counter = [0]
    parsed = _parse_token_list(pattern, tokens, counter)
    return Node(RE, pattern, 0, len(pattern), 0, 0, True, False, -1, [parsed])


# --------------------------------------------------------------------------- #
# Classifier: optimizer.mojo:103-846
def _classify_quantifier(ast: Node) -> int:
    """optimizer.mojo:320-349."""
    if ast.min == 1 and ast.max == 1:
        return SIMPLE
    if (ast.min == 0 and ast.max == -1) or (ast.min == 1 and ast.max == -1):
        return SIMPLE
    if ast.min == 0 and ast.max == 1:
        return SIMPLE
    if ast.max != -1 and ast.max - ast.min <= 10:
        return SIMPLE
    if ast.max != -1 and ast.max - ast.min <= 100:
        return MEDIUM
    return COMPLEX


def _opt_is_multi_char_class_sequence(ast: Node) -> bool:
    """optimizer.mojo:512-555."""
    if ast.type != GROUP or ast.get_children_len() < 2:
        return False
    cc = 0
    for e in ast.children:
        if e.type in (RANGE, DIGIT, WORD, SPACE, WILDCARD):
            cc += 1
        elif e.type == ELEMENT and e.min == 1 and e.max == 1:
            pass
        else:
            return False
    return cc >= 2


def _is_all_literal_branches(ast: Node) -> bool:
    """optimizer.mojo:492-510."""
    if ast.type == GROUP:
        if ast.get_children_len() == 1:
            return _is_all_literal_branches(ast.get_child(0))
        return all(c.type == ELEMENT for c in ast.children)
    if ast.type == OR:
        return all(_is_all_literal_branches(c) for c in ast.children)
    return ast.type == ELEMENT


def _is_simple_quantified_group(ast: Node) -> bool:
    """optimizer.mojo:557-586."""
    if not ((ast.min == 0 and ast.max == 1) or (ast.min == 0 and ast.max == -1)
            or (ast.min == 1 and ast.max == -1)):
        return False
    return all(c.type == ELEMENT and c.min == 1 and c.max == 1 for c in ast.children)


def _extract_literal_branches_strict(node: Node, branches: List[bytes]) -> bool:
    """optimizer.mojo:648-673 / 743-768, dfa.mojo:3462-3486, 3565-3589 (identical bodies)."""
    if node.type == OR:
        return (_extract_literal_branches_strict(node.get_child(0), branches)
                and _extract_literal_branches_strict(node.get_child(1), branches))
    if node.type == GROUP:
        text = b""
        for e in node.children:
            if e.type != ELEMENT:
                return False
            text += e.get_value() or b""
        branches.append(text)
        return True
    return False


def _common_prefix(branches: List[bytes]) -> bytes:
    """optimizer.mojo:675-710 / dfa.mojo:1465-1498, 3489-3522."""
    if not branches:
        return b""
    if len(branches) == 1:
        return branches[0]
    m = min(len(b) for b in branches)
    out = b""
    for pos in range(m):
        c = branches[0][pos]
        if all(b[pos] == c for b in branches[1:]):
            out += bytes([c])
        else:
            break
    return out


def _is_quantified_alternation_group_in_optimizer(ast: Node) -> bool:
    """optimizer.mojo:712-741."""
    if ast.min == 1 and ast.max == 1:
        return False
    if ast.get_children_len() != 1:
        return False
    orn = ast.get_child(0)
    if orn.type != OR:
        return False
    return _extract_literal_branches_strict(orn, [])


def _group_contains_or(ast: Node) -> bool:
    """optimizer.mojo:608-619."""
    for c in ast.children:
        if c.type == OR:
            return True
        if c.type == GROUP and _group_contains_or(c):
            return True
    return False


def _has_nested_alternation(ast: Node) -> bool:
    """optimizer.mojo:588-606."""
    for c in ast.children:
        if c.type == OR:
            if _has_nested_alternation(c):
                return True
        elif c.type == GROUP:
            if _group_contains_or(c):
                return True
    return False


def _is_simple_dfa_node(ast: Node) -> bool:
    """optimizer.mojo:827-845."""
    if ast.type in (ELEMENT, RANGE, DIGIT, WORD, SPACE, WILDCARD):
        return ast.max <= 10 or ast.max == -1
    return ast.type in (START, END)


def _is_dfa_compatible_branch(ast: Node) -> bool:
    """optimizer.mojo:794-825."""
    if ast.type == ELEMENT:
        return True
    if ast.type in (RANGE, DIGIT, WORD, SPACE):
        return True
    if ast.type == GROUP:
        if ast.get_children_len() <= 4:
            return all(_is_simple_dfa_node(c) for c in ast.children)
    elif ast.type == OR:
        if ast.get_children_len() <= 4:
            return all(_is_dfa_compatible_branch(c) for c in ast.children)
    return False


def _is_literal_heavy_alternation(ast: Node) -> bool:
    """optimizer.mojo:770-792."""
    if ast.type != OR:
        return False
    total = ast.get_children_len()
    ok = sum(1 for c in ast.children if _is_dfa_compatible_branch(c))
    return ok * 5 >= total * 4


def _is_common_prefix_alternation_in_tree(ast: Node) -> bool:
    """optimizer.mojo:621-646."""
    branches: List[bytes] = []
    if not _extract_literal_branches_strict(ast, branches):
        return False
    if len(branches) < 2:
        return False
    return len(_common_prefix(branches)) >= 2


def _analyze_alternation(ast: Node, depth: int) -> int:
    """optimizer.mojo:351-401."""
    if depth > 2:
        if _is_literal_heavy_alternation(ast):
            return MEDIUM
        if depth <= 4 and _is_common_prefix_alternation_in_tree(ast):
            return SIMPLE
        return COMPLEX
    mx = SIMPLE
    for c in ast.children:
        bc = _analyze_node(c, depth + 1)
        if bc == COMPLEX:
            return COMPLEX
        if bc == MEDIUM:
            mx = MEDIUM
    if mx == SIMPLE and ast.get_children_len() <= 8 and not _has_nested_alternation(ast):
        return SIMPLE
    return MEDIUM


def _analyze_group(ast: Node, depth: int) -> int:
    """optimizer.mojo:403-490."""
    if depth > 4:
        return COMPLEX
    qc = _classify_quantifier(ast)
    if qc == COMPLEX:
        return COMPLEX
    if ast.min != 1 or ast.max != 1:
        if _is_simple_quantified_group(ast):
            pass
        elif _is_quantified_alternation_group_in_optimizer(ast):
            return SIMPLE
        else:
            return MEDIUM
    if ast.get_children_len() == 1:
        only = ast.get_child(0)
        if only.type in (OR, GROUP) and _is_all_literal_branches(only):
            return SIMPLE
    mx = SIMPLE
    for c in ast.children:
        cc = _analyze_node(c, depth + 1)
        if cc == COMPLEX:
            return COMPLEX
        if cc == MEDIUM:
            mx = MEDIUM
    if mx == SIMPLE and qc == SIMPLE:
        all_literal = True
        for c in ast.children:
            is_lit = (c.type == ELEMENT and c.min == c.max
                      and 1 <= c.min <= MAX_LITERAL_QUANT_REPETITIONS)
            if not (is_lit or c.type in (START, END)):
                all_literal = False
                break
        if all_literal and ast.get_children_len() <= 20:
            return SIMPLE
        if ast.get_children_len() <= 5:
            return SIMPLE
        return MEDIUM
    return MEDIUM


def _analyze_node(ast: Node, depth: int) -> int:
    """optimizer.mojo:260-318."""
    if ast.type == RE:
        if ast.get_children_len() == 0:
            return SIMPLE
        return _analyze_node(ast.get_child(0), depth)
    if ast.type in (ELEMENT, WILDCARD, SPACE, DIGIT, WORD, RANGE):
        return _classify_quantifier(ast)
    if ast.type in (START, END):
        return SIMPLE
    if ast.type == OR:
        return _analyze_alternation(ast, depth)
    if ast.type == GROUP:
        if _opt_is_multi_char_class_sequence(ast):
            return SIMPLE
        return _analyze_group(ast, depth)
    return COMPLEX


def classify(ast: Node) -> int:
    """optimizer.mojo:110-119."""
    return _analyze_node(ast, 0)


def _count_simd_nodes(ast: Node) -> int:
    """optimizer.mojo:234-258."""
    count = 0
    if ast.type in (RANGE, DIGIT, WORD, SPACE):
        count += 2 if (ast.min > 1 or ast.max == -1) else 1
    elif ast.type in (GROUP, RE, OR):
        for c in ast.children:
            count += _count_simd_nodes(c)
    return count


def should_use_pure_dfa(ast: Node) -> bool:
    """optimizer.mojo:174-201."""
    if classify(ast) != SIMPLE:
        return False
    return _count_simd_nodes(ast) <= 1


def _is_literal_sequence(ast: Node) -> bool:
    """optimizer.mojo:866-900."""
    if ast.type == ELEMENT:
        if ast.min == 1 and ast.max == 1:
            return True
        return ast.min == ast.max and 1 <= ast.min <= MAX_LITERAL_QUANT_REPETITIONS
    if ast.type in (START, END):
        return True
    if ast.type == GROUP:
        for c in ast.children:
            if c.type == GROUP:
                return False
            if not _is_literal_sequence(c):
                return False
        return True
    return False


def is_literal_pattern(ast: Node) -> bool:
    """optimizer.mojo:848-863."""
    if ast.type != RE:
        return False
    if not ast.has_children():
        return True
    return _is_literal_sequence(ast.get_child(0))


def _extract_literal_chars(ast: Node) -> bytes:
    """optimizer.mojo:921-953."""
    if ast.type == ELEMENT:
        v = ast.get_value()
        if not v:
            return b""
        return v if ast.min <= 1 else v * ast.min
    if ast.type == GROUP:
        return b"".join(_extract_literal_chars(c) for c in ast.children)
    return b""


def get_literal_string(ast: Node) -> bytes:
    """optimizer.mojo:903-918."""
    if ast.type == RE and ast.get_children_len() > 0:
        return _extract_literal_chars(ast.get_child(0))
    return b""


def _check_anchors_recursive(ast: Node) -> Tuple[bool, bool]:
    """optimizer.mojo:975-999."""
    if ast.type == START:
        return (True, False)
    if ast.type == END:
        return (False, True)
    if ast.type == GROUP:
        hs = he = False
        for c in ast.children:
            s, e = _check_anchors_recursive(c)
            hs, he = hs or s, he or e
        return (hs, he)
    return (False, False)


def pattern_has_anchors(ast: Node) -> Tuple[bool, bool]:
    """optimizer.mojo:956-972."""
    if ast.type == RE and ast.has_children():
        return _check_anchors_recursive(ast.get_child(0))
    return (False, False)


# --------------------------------------------------------------------------- #
# Literal extraction for the prefilter: literal_optimizer.mojo:24-238, matcher.mojo:99-108
@dataclass
class _Lit:
    literal: bytes
    start_offset: int
    is_prefix: bool
    is_required: bool


def _extract_sequence(group: Node, start_offset: int, is_required: bool, at_start: bool,
                      out: List[_Lit]) -> None:
    """literal_optimizer.mojo:333-394."""
    cur = b""
    cur_off = start_offset
    seq_at_start = at_start
    for child in group.children:
        if child.type == ELEMENT and child.min == 1 and child.max == 1 and child.get_value():
            cur += child.get_value()
        else:
            if cur:
                out.append(_Lit(cur, cur_off, seq_at_start, is_required))
                cur_off += len(cur)
                seq_at_start = False
                cur = b""
            if child.type in (START, END):
                continue
            seq_at_start = False
            if child.min > 0:
                cur_off += 1
    if cur:
        out.append(_Lit(cur, cur_off, seq_at_start, is_required))


def _get_prefix_literal(node: Node) -> bytes:
    """literal_optimizer.mojo:430-447."""
    if node.type == ELEMENT and node.min >= 1 and node.max >= 1 and node.get_value():
        return node.get_value()
    if node.type == GROUP and node.min >= 1:
        lits: List[_Lit] = []
        _extract_sequence(node, 0, True, True, lits)
        if lits:
            return lits[0].literal
    return b""


def _collect_or_prefixes(node: Node, prefixes: List[bytes]) -> None:
    """literal_optimizer.mojo:416-427."""
    if node.type != OR:
        p = _get_prefix_literal(node)
        if p:
            prefixes.append(p)
        return
    for c in node.children:
        _collect_or_prefixes(c, prefixes)


def _find_common_prefix_simple(or_node: Node) -> bytes:
    """literal_optimizer.mojo:397-413."""
    prefixes: List[bytes] = []
    _collect_or_prefixes(or_node, prefixes)
    if len(prefixes) < 2:
        return b""
    common = prefixes[0]
    for p in prefixes[1:]:
        k = 0
        while k < min(len(common), len(p)) and common[k] == p[k]:
            k += 1
        common = common[:k]
        if not common:
            return b""
    return common


def _extract_from_node(node: Node, out: List[_Lit], offset: int, is_required: bool,
                       at_start: bool) -> None:
    """literal_optimizer.mojo:241-330."""
    if node.type == ELEMENT:
        if node.min >= 1 and node.get_value():
            if node.max == 1 or node.max == -1:
                out.append(_Lit(node.get_value(), offset, at_start, is_required))
    elif node.type == GROUP:
        if node.min >= 1:
            if node.get_children_len() == 1:
                child = node.get_child(0)
                if child.type in (GROUP, OR):
                    _extract_from_node(child, out, offset, is_required, at_start)
                    return
            _extract_sequence(node, offset, is_required, at_start, out)
    elif node.type == OR:
        cp = _find_common_prefix_simple(node)
        if cp:
            out.append(_Lit(cp, offset, at_start, True))
        for c in node.children:
            _extract_from_node(c, out, offset, False, at_start)


def best_literal(ast: Node) -> Optional[_Lit]:
    """literal_optimizer.mojo:217-238 + select_best :166-206 (first max score wins; a
    score of 0 can never beat the initial best_score=0, so index 0 is kept)."""
    lits: List[_Lit] = []
    if ast.type == RE and ast.has_children():
        _extract_from_node(ast.get_child(0), lits, 0, True, True)
    elif ast.type == GROUP:
        _extract_from_node(ast, lits, 0, True, True)
    if not lits:
        return None
    best_idx, best_score = 0, 0
    for i, lit in enumerate(lits):
        score = (1000 if lit.is_required else 0) + len(lit.literal) * 10
        score += 100 if lit.is_prefix else 0
        score += lit.start_offset
        if score > best_score:
            best_score, best_idx = score, i
    return lits[best_idx]


def _has_literal_prefix_node(node: Node) -> bool:
    """literal_optimizer.mojo:479-497."""
    if node.type == START:
        return False
    if node.type == ELEMENT:
        return node.min >= 1 and node.max >= 1
    if node.type == GROUP:
        if node.min >= 1 and node.get_children_len() > 0:
            first = node.get_child(0)
            if first.type == START and node.get_children_len() > 1:
                return _has_literal_prefix_node(node.get_child(1))
            return _has_literal_prefix_node(first)
    return False


def has_literal_prefix(ast: Node) -> bool:
    """literal_optimizer.mojo:463-476."""
    if ast.type != RE or not ast.has_children():
        return False
    return _has_literal_prefix_node(ast.get_child(0))


def check_ast_for_anchors(ast: Node) -> bool:
    """matcher.mojo:168-178."""
    if ast.type in (START, END):
        return True
    if ast.type in (GROUP, RE):
        return any(check_ast_for_anchors(c) for c in ast.children)
    return False


def is_simple_pattern_skip_prefilter(pattern: bytes) -> bool:
    """matcher.mojo:447-532."""
    n = len(pattern)
    if n <= 4:
        return True
    has_q = has_alt = has_anch = has_wild = has_cc = False
    for i, c in enumerate(pattern):
        if c in (ord("*"), ord("+"), ord("?")):
            has_q = True
        elif c == ord("|"):
            has_alt = True
        elif c == ord("^") and i == 0:
            has_anch = True
        elif c == ord("$") and i == n - 1:
            has_anch = True
        elif c == ord("."):
            has_wild = True
        elif c == ord("["):
            has_cc = True
    if has_q and not has_alt and not has_wild:
        return True
    if has_alt and n <= 10 and not has_wild:
        if pattern.count(b"|") >= n // 3:
            return True
    if has_anch and n <= 10:
        return True
    if has_wild and n >= 8:
        return False
    if has_alt and n >= 12:
        return False
    if n >= 15:
        return False
    return True


# --------------------------------------------------------------------------- #
# CharacterClassSIMD: simd_ops.mojo:261-420, 62-86, 1177-1216
class CharClass:
    """Lookup table + the range / nibble metadata that decides which scan loop the
    reference runs (simd_ops.mojo:293-401)."""

    def __init__(self, chars: bytes):
        self.lookup = [0] * 256
        for b in chars:
            self.lookup[b] = 1
        # _detect_ranges, simd_ops.mojo:364-401
        starts, ends, count, in_range = [], [], 0, False
        for c in range(256):
            if self.lookup[c]:
                if not in_range:
                    starts.append(c)
                    in_range = True
            elif in_range:
                ends.append(c - 1)
                count += 1
                in_range = False
        if in_range:
            ends.append(255)
            count += 1
        self.num_ranges = 0 if count > 3 else count
        # build_nibble_tables, simd_ops.mojo:62-86 (bucket bits wrap at 8 -> false
        # positives for classes with more than 8 members)
        lo = [0] * 16
        hi = [0] * 16
        bucket = 0
        for c in range(256):
            if self.lookup[c]:
                bit = 1 << (bucket & 7)
                lo[c & 15] |= bit
                hi[(c >> 4) & 15] |= bit
                bucket += 1
        # candidate table of find_first_nibble_match's SIMD body (simd_ops.mojo:565-648):
        # exact for 1..3 ranges, nibble-product otherwise.
        if self.num_ranges in (1, 2, 3):
            self.simd_candidates = list(self.lookup)
        else:
            self.simd_candidates = [1 if (lo[c & 15] & hi[c >> 4]) else 0 for c in range(256)]


def get_character_class_matcher(char_class: bytes) -> CharClass:
    """simd_ops.mojo:1177-1216 -- the cached whitespace matcher carries '\\v'
    (simd_ops.mojo:879-883) even when asked for the 5-char DFA set."""
    if char_class in (b" \t\n\r\x0c", b" \t\n\r\x0c\x0b"):
        return CharClass(b" \t\n\r\x0c\x0b")
    return CharClass(char_class)


# --------------------------------------------------------------------------- #
# DFA tables: dfa.mojo:215-254, 257-306
@dataclass
class DFA:
    trans: List[List[int]] = field(default_factory=list)  # n_states x 256, -1 = none
    accepting: List[bool] = field(default_factory=list)
    start_state: int = 0
    has_start_anchor: bool = False
    has_end_anchor: bool = False
    is_pure_literal: bool = False
    literal_pattern: bytes = b""
    matcher: Optional[CharClass] = None  # _simd_char_matcher / _has_simd_matcher
    simd_scan_eligible: bool = False

    def new_state(self, accepting: bool = False) -> int:
        self.trans.append([-1] * 256)
        self.accepting.append(accepting)
        return len(self.trans) - 1

    def add_transition(self, s: int, c: int, t: int) -> None:
        """dfa.mojo:233-242 -- later writes overwrite earlier ones."""
        if 0 <= c < 256:
            self.trans[s][c] = t

    def add_class(self, frm: int, to: int, char_class: bytes, positive: bool) -> None:
        """dfa.mojo:1747-1803."""
        if frm >= len(self.trans):
            return
        if positive:
            for c in char_class:
                self.add_transition(frm, c, to)
        else:
            self.trans[frm] = [to] * 256
            for c in char_class:
                self.trans[frm][c] = -1

    def find_or_create_state(self, frm: int, c: int) -> int:
        """dfa.mojo:1245-1268."""
        t = self.trans[frm][c]
        if t != -1:
            return t
        ns = self.new_state()
        self.add_transition(frm, c, ns)
        return ns


def expand_character_range(range_str: bytes) -> bytes:
    """dfa.mojo:71-168.  The dispatcher always passes the *root* node type (RE), so the
    node-type fast path (:84-89) never fires from compile_dfa_pattern; the callers that
    do hold a DIGIT/WORD/SPACE node return their constant before calling this."""
    if not (range_str.startswith(b"[") and range_str.endswith(b"]")):
        return range_str
    if range_str == b"[a-z]":
        return LOWERCASE_LETTERS
    if range_str == b"[A-Z]":
        return UPPERCASE_LETTERS
    if range_str == b"[0-9]":
        return DIGITS
    if range_str == b"[a-zA-Z0-9]":
        return ALPHANUMERIC
    if range_str == b"[a-zA-Z]":
        return ALL_LETTERS
    inner = range_str[1:-1]
    if inner.startswith(b"^"):
        inner = inner[1:]
    if len(inner) == 3 and inner[1] == ord("-"):
        s, e = inner[0], inner[2]
        if s >= ord("a") and e <= ord("z"):
            return LOWERCASE_LETTERS[s - ord("a"): e - ord("a") + 1]
        elif s >= ord("a") and e <= ord("Z"):
            return UPPERCASE_LETTERS[s - ord("A"): e - ord("A") + 1]
        elif s >= ord("0") and e <= ord("9"):
            return DIGITS[s - ord("0"): e - ord("0") + 1]
    out = bytearray()
    i = 0
    while i < len(inner):
        if i + 2 < len(inner) and inner[i + 1] == ord("-"):
            out.extend(range(inner[i], inner[i + 2] + 1)) if inner[i] <= inner[i + 2] else None
            i += 3
        else:
            out.append(inner[i])
            i += 1
    return bytes(b for b in out if b < 256)


@dataclass
class SeqElement:
    """dfa.mojo:171-196."""
    char_class: bytes
    min_matches: int
    max_matches: int
    positive_logic: bool = True
    alternation_branches: List[bytes] = field(default_factory=list)


# ---- builders -------------------------------------------------------------- #
def compile_pattern(dfa: DFA, pattern: bytes, has_start: bool, has_end: bool) -> None:
    """dfa.mojo:308-352 (literal chain)."""
    dfa.has_start_anchor, dfa.has_end_anchor = has_start, has_end
    dfa.literal_pattern = pattern
    if len(pattern) == 0:
        dfa.new_state(True)  # _create_accepting_state, dfa.mojo:1500-1505
        return
    if not has_start and not has_end:
        dfa.is_pure_literal = True
    for i, c in enumerate(pattern):
        s = dfa.new_state()
        dfa.add_transition(s, c, i + 1)
    dfa.new_state(True)
    dfa.start_state = 0


def compile_character_class_with_logic(dfa: DFA, char_class: bytes, mn: int, mx: int,
                                       positive: bool) -> None:
    """dfa.mojo:375-496."""
    if mn >= 0 and positive:
        dfa.matcher = get_character_class_matcher(char_class)
        dfa.simd_scan_eligible = mx == -1
    if mn == 0:
        dfa.new_state(True)
        dfa.new_state(True)
        dfa.add_class(0, 1, char_class, positive)
        if mx == -1 or mx > 1:
            dfa.add_class(1, 1, char_class, positive)
    elif mn == 1:
        dfa.new_state(False)
        dfa.new_state(True)
        dfa.add_class(0, 1, char_class, positive)
        if mx == -1:
            dfa.add_class(1, 1, char_class, positive)
        elif mx > 1:
            for k in range(2, mx + 1):
                dfa.new_state(True)
                dfa.add_class(k - 1, k, char_class, positive)
    else:
        for k in range(mn + 1):
            dfa.new_state(k >= mn)
            if k > 0:
                dfa.add_class(k - 1, k, char_class, positive)
        if mx == -1:
            last = len(dfa.trans) - 1
            dfa.add_class(last, last, char_class, positive)
        elif mx > mn:
            for k in range(mn + 1, mx + 1):
                dfa.new_state(True)
                dfa.add_class(k - 1, k, char_class, positive)
    dfa.start_state = 0


def compile_sequential_pattern(dfa: DFA, elements: List[SeqElement], hs: bool, he: bool) -> None:
    """dfa.mojo:498-605."""
    dfa.has_start_anchor, dfa.has_end_anchor = hs, he
    if not elements:
        dfa.new_state(True)
        return
    cur = 0
    for idx, el in enumerate(elements):
        is_last = idx == len(elements) - 1
        if el.min_matches == 0:
            if idx == 0:
                dfa.new_state(not is_last)
                cur = 0
            ms = dfa.new_state(True)
            dfa.add_class(cur, ms, el.char_class, el.positive_logic)
            if el.max_matches == -1:
                dfa.add_class(ms, ms, el.char_class, el.positive_logic)
            cur = ms
        else:
            for k in range(el.min_matches):
                acc = (k >= el.min_matches - 1) and is_last
                si = dfa.new_state(acc)
                if k == 0:
                    dfa.add_class(cur, si, el.char_class, el.positive_logic)
                else:
                    dfa.add_class(si - 1, si, el.char_class, el.positive_logic)
                cur = si
            if el.max_matches == -1:
                dfa.add_class(cur, cur, el.char_class, el.positive_logic)
            elif el.max_matches > el.min_matches:
                for _ in range(el.max_matches - el.min_matches):
                    si = dfa.new_state(is_last)
                    dfa.add_class(cur, si, el.char_class, el.positive_logic)
                    cur = si
    dfa.start_state = 0


def compile_multi_character_class_sequence(dfa: DFA, elements: List[SeqElement], hs: bool,
                                           he: bool) -> None:
    """dfa.mojo:607-871."""
    dfa.has_start_anchor, dfa.has_end_anchor = hs, he
    if not elements:
        dfa.new_state(True)
        return
    if len(elements[0].alternation_branches) == 0 and len(elements[0].char_class) > 0:
        dfa.matcher = get_character_class_matcher(elements[0].char_class)
        # not scan eligible (dfa.mojo:643-644)
    cur = 0
    n_el = len(elements)
    for idx, el in enumerate(elements):
        is_last = idx == n_el - 1
        all_rest_optional = all(elements[i].min_matches <= 0 for i in range(idx + 1, n_el))
        if el.alternation_branches:  # dfa.mojo:660-697
            if idx == 0:
                dfa.new_state(False)
                cur = 0
            end_state = dfa.new_state(is_last or all_rest_optional)
            for branch in el.alternation_branches:
                prev = cur
                for ci, c in enumerate(branch):
                    if ci == len(branch) - 1:
                        dfa.add_transition(prev, c, end_state)
                    else:
                        mid = dfa.new_state(False)
                        dfa.add_transition(prev, c, mid)
                        prev = mid
            cur = end_state
            continue
        if el.min_matches == 0:  # dfa.mojo:699-744
            if idx == 0:
                dfa.new_state(all_rest_optional)
                cur = 0
            ms = dfa.new_state(is_last or all_rest_optional)
            dfa.add_class(cur, ms, el.char_class, el.positive_logic)
            if el.max_matches == -1:
                dfa.add_class(ms, ms, el.char_class, el.positive_logic)
            cur = 0 if idx == 0 else ms
        elif el.min_matches == 1:  # dfa.mojo:746-807
            if idx == 0:
                dfa.new_state(False)
                cur = 0
            ms = dfa.new_state(is_last or all_rest_optional)
            dfa.add_class(cur, ms, el.char_class, el.positive_logic)
            if idx == 1 and cur == 0 and elements[0].min_matches == 0:
                dfa.add_class(1, ms, el.char_class, el.positive_logic)
            if el.max_matches == -1:
                dfa.add_class(ms, ms, el.char_class, el.positive_logic)
            elif el.max_matches > 1:
                for k in range(2, el.max_matches + 1):
                    extra = dfa.new_state(is_last)
                    dfa.add_class(ms + k - 2, extra, el.char_class, el.positive_logic)
            cur = ms
        else:  # dfa.mojo:809-869
            if idx == 0:
                dfa.new_state(False)
                cur = 0
            for k in range(el.min_matches):
                acc = (k >= el.min_matches - 1) and is_last
                si = dfa.new_state(acc)
                if k > 0:
                    dfa.add_class(si - 1, si, el.char_class, el.positive_logic)
                else:
                    dfa.add_class(cur, si, el.char_class, el.positive_logic)
                cur = si
            if el.max_matches == -1:
                dfa.add_class(cur, cur, el.char_class, el.positive_logic)
            elif el.max_matches > el.min_matches:
                for _ in range(el.min_matches + 1, el.max_matches + 1):
                    oi = dfa.new_state(is_last)
                    dfa.add_class(cur, oi, el.char_class, el.positive_logic)
                    cur = oi
    dfa.start_state = 0


def _extract_branch_text(branch: Node) -> bytes:
    """dfa.mojo:3172-3194."""
    if branch.type == ELEMENT and branch.get_value():
        return branch.get_value()
    if branch.type == GROUP:
        return b"".join(c.get_value() for c in branch.children
                        if c.type == ELEMENT and c.get_value())
    return b""


def _collect_all_alternation_branches(or_node: Node) -> List[bytes]:
    """dfa.mojo:3265-3306."""
    out: List[bytes] = []
    for branch in or_node.children:
        if branch.type == OR:
            out.extend(_collect_all_alternation_branches(branch))
        elif branch.type == GROUP:
            inner = branch
            while inner.type == GROUP and inner.get_children_len() == 1:
                inner = inner.get_child(0)
            if inner.type == OR:
                out.extend(_collect_all_alternation_branches(inner))
            else:
                t = _extract_branch_text(branch)
                if t:
                    out.append(t)
        else:
            t = _extract_branch_text(branch)
            if t:
                out.append(t)
    return out


def _find_or_node(ast: Node) -> Optional[Node]:
    """dfa.mojo:3146-3169."""
    if ast.type == OR:
        return ast
    for c in ast.children:
        f = _find_or_node(c)
        if f is not None:
            return f
    return None


def compile_alternation(dfa: DFA, ast: Node) -> None:
    """dfa.mojo:873-928 -- trie with one shared accepting state."""
    dfa.trans.clear()
    dfa.accepting.clear()
    dfa.new_state(False)
    dfa.start_state = 0
    or_node = _find_or_node(ast)
    if or_node is None:
        raise Unsupported("No OR node found in alternation pattern")
    acc = dfa.new_state(True)
    for text in _collect_all_alternation_branches(or_node):
        if not text:
            continue
        cur = 0
        for j, c in enumerate(text):
            if j == len(text) - 1:
                dfa.add_transition(cur, c, acc)
            else:
                cur = dfa.find_or_create_state(cur, c)


def _group_text(group: Node) -> bytes:
    """dfa.mojo:3245-3262 (get_value() of an empty slice would trap in Mojo; b'' here)."""
    return b"".join((c.get_value() or b"") for c in group.children if c.type == ELEMENT)


def _compile_looping_literal(dfa: DFA, text: bytes, mn: int, mx: int, acc: int,
                             optional_via_epsilon: bool) -> None:
    """Shared body of compile_quantified_group (dfa.mojo:982-1074) and
    compile_simple_quantifier (:1147-1243).  The two differ only in how `?` encodes the
    zero-repeat case: the group variant adds a transition on byte 0 to the accepting
    state (:989), the simple variant marks the start state accepting (:1154-1156)."""
    if mn == 0 and mx == 1:
        if optional_via_epsilon:
            dfa.add_transition(0, 0, acc)
        else:
            dfa.accepting[0] = True
        cur = 0
        for i, c in enumerate(text):
            if i == len(text) - 1:
                dfa.add_transition(cur, c, acc)
            else:
                ns = dfa.new_state()
                dfa.add_transition(cur, c, ns)
                cur = ns
    elif mn == 0 and mx == -1:
        dfa.accepting[0] = True
        cur = 0
        for i, c in enumerate(text):
            if i == len(text) - 1:
                dfa.add_transition(cur, c, 0)
            else:
                ns = dfa.new_state()
                dfa.add_transition(cur, c, ns)
                cur = ns
    elif mn == 1 and mx == -1:
        cur = 0
        for i, c in enumerate(text):
            if i == len(text) - 1:
                loop = dfa.new_state(True)
                dfa.add_transition(cur, c, loop)
                dfa.add_transition(loop, text[0], 1 if len(text) > 1 else loop)
            else:
                ns = dfa.new_state()
                dfa.add_transition(cur, c, ns)
                cur = ns
    else:
        raise Unsupported("Unsupported quantifier range")


def compile_quantified_group(dfa: DFA, ast: Node) -> None:
    """dfa.mojo:930-980."""
    dfa.trans.clear()
    dfa.accepting.clear()
    dfa.new_state(False)
    inner = ast.get_child(0).get_child(0)
    text = _group_text(inner)
    if not text:
        raise Unsupported("Empty quantified group")
    acc = dfa.new_state(True)
    _compile_looping_literal(dfa, text, inner.min, inner.max, acc, True)


def compile_simple_quantifier(dfa: DFA, ast: Node) -> None:
    """dfa.mojo:1076-1145 -- all ELEMENT bytes are concatenated and the *last seen*
    quantifier is applied to the whole string."""
    dfa.trans.clear()
    dfa.accepting.clear()
    dfa.new_state(False)
    group = ast.get_child(0)
    qmin, qmax = 1, 1
    text = b""
    for el in group.children:
        if el.min == 0 and el.max == -1:
            qmin, qmax = 0, -1
        elif el.min == 1 and el.max == -1:
            qmin, qmax = 1, -1
        elif el.min == 0 and el.max == 1:
            qmin, qmax = 0, 1
        text += el.get_value() or b""
    if not text:
        raise Unsupported("Empty quantifier pattern")
    acc = dfa.new_state(True)
    _compile_looping_literal(dfa, text, qmin, qmax, acc, False)


def compile_wildcard_quantifier(dfa: DFA, ast: Node) -> None:
    """dfa.mojo:1270-1361."""
    dfa.trans.clear()
    dfa.accepting.clear()
    dfa.new_state(False)
    w = ast.get_child(0).get_child(0)
    acc = dfa.new_state(True)
    not_nl = [i for i in range(256) if i != 10]
    if w.min == 0 and w.max == 1:
        dfa.accepting[0] = True
        for i in not_nl:
            dfa.add_transition(0, i, acc)
    elif w.min == 0 and w.max == -1:
        dfa.accepting[0] = True
        for i in not_nl:
            dfa.add_transition(0, i, 0)
    elif w.min == 1 and w.max == -1:
        loop = dfa.new_state(True)
        for i in not_nl:
            dfa.add_transition(0, i, loop)
            dfa.add_transition(loop, i, loop)
    elif w.min == 1 and w.max == 1:
        for i in not_nl:
            dfa.add_transition(0, i, acc)
    else:
        raise Unsupported("Unsupported quantifier type for wildcard quantifier")


def _lenient_branches(node: Node, branches: List[bytes]) -> None:
    """dfa.mojo:1396-1418 / 1556-1581 (non-ELEMENT children are skipped, not rejected)."""
    if node.type == OR:
        _lenient_branches(node.get_child(0), branches)
        _lenient_branches(node.get_child(1), branches)
    elif node.type == GROUP:
        branches.append(b"".join((e.get_value() or b"") for e in node.children
                                 if e.type == ELEMENT))


def compile_common_prefix_alternation(dfa: DFA, ast: Node) -> None:
    """dfa.mojo:1363-1463."""
    dfa.trans.clear()
    dfa.accepting.clear()
    dfa.new_state(False)
    or_node = ast.get_child(0).get_child(0).get_child(0)
    branches: List[bytes] = []
    _lenient_branches(or_node, branches)
    if not branches:
        return
    prefix = _common_prefix(branches)
    cur = 0
    for c in prefix:
        cur = dfa.find_or_create_state(cur, c)
    for b in branches:
        if len(b) == len(prefix):
            dfa.accepting[cur] = True
        else:
            sc = cur
            suffix = b[len(prefix):]
            for j, c in enumerate(suffix):
                if j == len(suffix) - 1:
                    t = dfa.find_or_create_state(sc, c)
                    dfa.accepting[t] = True
                else:
                    sc = dfa.find_or_create_state(sc, c)


def compile_quantified_alternation_group(dfa: DFA, ast: Node) -> None:
    """dfa.mojo:1507-1686."""
    dfa.trans.clear()
    dfa.accepting.clear()
    dfa.new_state(False)
    qg = ast.get_child(0).get_child(0)
    or_node = qg.get_child(0)
    branches: List[bytes] = []
    _lenient_branches(or_node, branches)
    if qg.min == 0 and qg.max == 1:
        dfa.accepting[0] = True
        acc = dfa.new_state(True)
        for b in branches:
            cur = 0
            for j, c in enumerate(b):
                if j == len(b) - 1:
                    dfa.add_transition(cur, c, acc)
                else:
                    cur = dfa.find_or_create_state(cur, c)
    elif qg.min == 0 and qg.max == -1:
        dfa.accepting[0] = True
        for b in branches:
            cur = 0
            for j, c in enumerate(b):
                if j == len(b) - 1:
                    dfa.add_transition(cur, c, 0)
                else:
                    cur = dfa.find_or_create_state(cur, c)
    elif qg.min == 1 and qg.max == -1:
        loop = dfa.new_state(True)
        for origin in (0, loop):
            for b in branches:
                cur = origin
                for j, c in enumerate(b):
                    if j == len(b) - 1:
                        dfa.add_transition(cur, c, loop)
                    else:
                        cur = dfa.find_or_create_state(cur, c)
    else:
        raise Unsupported("Unsupported quantifier type for quantified alternation group")


# ---- shape predicates / extractors ----------------------------------------- #
def _element_to_char_class(el: Node) -> bytes:
    """dfa.mojo:2885-2908."""
    if el.type == DIGIT:
        return DIGITS
    if el.type == WORD:
        return WORD_CHARS
    if el.type == RANGE:
        return expand_character_range(el.get_value() or b"")
    if el.type == SPACE:
        return SPACE_CLASS
    if el.type == WILDCARD:
        return ALL_EXCEPT_NEWLINE
    if el.type == ELEMENT and el.get_value():
        return el.get_value()
    return b""


def _is_char_class_group(node: Node) -> bool:
    """dfa.mojo:2711-2733."""
    if node.type != GROUP:
        return False
    has_cc = False
    for c in node.children:
        if not _element_to_char_class(c):
            return False
        if c.type != ELEMENT:
            has_cc = True
    return has_cc


def _is_literal_alternation_group(node: Node) -> bool:
    """dfa.mojo:2736-2771."""
    if node.type != GROUP or node.get_children_len() != 1:
        return False
    child = node.get_child(0)
    if child.type != OR:
        return False
    has_branch = False
    stack = [child]
    while stack:
        cur = stack.pop()
        if cur.type == GROUP:
            if any(c.type != ELEMENT for c in cur.children):
                return False
            has_branch = True
        elif cur.type == OR:
            if cur.get_children_len() != 2:
                return False
            stack.append(cur.get_child(1))
            stack.append(cur.get_child(0))
        else:
            return False
    return has_branch


def _collect_alternation_branches(node: Node) -> List[bytes]:
    """dfa.mojo:2774-2812."""
    branches: List[bytes] = []
    stack: List[Node] = []
    if node.get_children_len() > 0:
        stack.append(node.get_child(0))
    while stack:
        cur = stack.pop()
        if cur.type == GROUP:
            branches.append(b"".join(c.get_value() for c in cur.children
                                     if c.type == ELEMENT and c.get_value()))
        elif cur.type == OR:
            if cur.get_children_len() >= 2:
                stack.append(cur.get_child(1))
            if cur.get_children_len() >= 1:
                stack.append(cur.get_child(0))
    return branches


def _is_multi_character_class_sequence(ast: Node) -> bool:
    """dfa.mojo:2815-2882."""
    if ast.type != RE or ast.get_children_len() != 1:
        return False
    child = ast.get_child(0)
    if child.type != GROUP or child.get_children_len() < 2:
        return False
    cc = 0
    for e in child.children:
        if e.type in (RANGE, DIGIT, WORD, SPACE, WILDCARD):
            cc += 1
        elif e.type == ELEMENT and e.min == 1 and e.max == 1:
            pass
        elif e.type == GROUP and _is_literal_alternation_group(e):
            pass
        elif e.type == GROUP and _is_char_class_group(e):
            cc += 1
        else:
            return False
    return cc >= 2


def _extract_multi_class_sequence_info(ast: Node) -> Tuple[List[SeqElement], bool, bool]:
    """dfa.mojo:2911-2970."""
    hs, he = pattern_has_anchors(ast)
    els: List[SeqElement] = []
    if ast.type == RE and ast.get_children_len() == 1:
        child = ast.get_child(0)
        if child.type == GROUP:
            for e in child.children:
                if e.type == GROUP and _is_literal_alternation_group(e):
                    se = SeqElement(b"", 1, 1, True)
                    se.alternation_branches = _collect_alternation_branches(e)
                    els.append(se)
                elif e.type == GROUP and _is_char_class_group(e):
                    for sub in e.children:
                        sc = _element_to_char_class(sub)
                        if sc:
                            els.append(SeqElement(sc, sub.min, sub.max, sub.positive_logic))
                else:
                    cc = _element_to_char_class(e)
                    if cc:
                        els.append(SeqElement(cc, e.min, e.max, e.positive_logic))
    return els, hs, he


def _is_simple_character_class_pattern(ast: Node) -> bool:
    """dfa.mojo:2499-2529."""
    if _is_multi_character_class_sequence(ast):
        return False
    if ast.type == RE and ast.get_children_len() == 1:
        child = ast.get_child(0)
        if child.type in (DIGIT, WORD, RANGE):
            return True
        if child.type == GROUP and child.get_children_len() == 1:
            return child.get_child(0).type in (DIGIT, WORD, RANGE)
    elif ast.type in (DIGIT, WORD, RANGE):
        return True
    return False


def _is_pure_anchor_pattern(ast: Node) -> bool:
    """dfa.mojo:2605-2629."""
    if ast.type in (START, END):
        return True
    if ast.type == RE:
        if not ast.has_children():
            return False
        return _is_pure_anchor_pattern(ast.get_child(0))
    if ast.type == GROUP:
        return all(_is_pure_anchor_pattern(c) for c in ast.children)
    return False


def _is_sequential_character_class_pattern(ast: Node) -> bool:
    """dfa.mojo:2632-2663."""
    if ast.type != RE or ast.get_children_len() != 1:
        return False
    child = ast.get_child(0)
    if child.type != GROUP:
        return False
    if any(e.type not in (RANGE, DIGIT, WORD) for e in child.children):
        return False
    return child.get_children_len() >= 2


def _is_mixed_sequential_pattern(ast: Node) -> bool:
    """dfa.mojo:2973-3015."""
    if ast.type != RE or ast.get_children_len() != 1:
        return False
    child = ast.get_child(0)
    if child.type != GROUP or child.get_children_len() < 3:
        return False
    has_cc = has_opt = False
    for e in child.children:
        if e.type in (RANGE, DIGIT, WORD):
            has_cc = True
        elif e.type == ELEMENT and e.min == 0 and e.max == 1:
            has_opt = True
    return has_cc and has_opt


def _group_contains_only_literals(group: Node) -> bool:
    """dfa.mojo:3100-3120."""
    return group.type == GROUP and all(c.type == ELEMENT for c in group.children)


def _is_simple_alternation_branches(ast: Node) -> bool:
    """dfa.mojo:3054-3097."""
    if ast.type != OR:
        return False
    for branch in ast.children:
        if branch.type == GROUP:
            if not _group_contains_only_literals(branch):
                inner = branch
                while inner.type == GROUP and inner.get_children_len() == 1:
                    inner = inner.get_child(0)
                if inner.type == OR:
                    if not _is_simple_alternation_branches(inner):
                        return False
                else:
                    return False
        elif branch.type == ELEMENT:
            continue
        elif branch.type == OR:
            if not _is_simple_alternation_branches(branch):
                return False
        else:
            return False
    return True


def _is_alternation_pattern(ast: Node) -> bool:
    """dfa.mojo:3034-3051 + _is_pure_alternation_pattern :3309-3342."""
    if ast.type == OR:
        return _is_simple_alternation_branches(ast)
    if ast.type != RE or ast.get_children_len() != 1:
        return False
    child = ast.get_child(0)
    if child.type == OR:
        return _is_simple_alternation_branches(child)
    node = child
    while node.type == GROUP and node.get_children_len() == 1:
        node = node.get_child(0)
    if node.type == OR:
        return _is_simple_alternation_branches(node)
    return False


def _is_quantified_group(ast: Node) -> bool:
    """dfa.mojo:3197-3242."""
    if ast.type == RE and ast.get_children_len() == 1:
        child = ast.get_child(0)
        if child.type == GROUP and child.get_children_len() == 1:
            g = child.get_child(0)
            if g.type == GROUP and (g.min != 1 or g.max != 1):
                return all(c.type == ELEMENT for c in g.children)
    return False


def _is_simple_quantifier_pattern(ast: Node) -> bool:
    """dfa.mojo:3345-3383."""
    if ast.type != RE or ast.get_children_len() != 1:
        return False
    group = ast.get_child(0)
    if group.type != GROUP or group.get_children_len() == 0:
        return False
    has_q = False
    for c in group.children:
        if c.type != ELEMENT:
            return False
        if ((c.min == 0 and c.max == -1) or (c.min == 1 and c.max == -1)
                or (c.min == 0 and c.max == 1)):
            has_q = True
    return has_q


def _is_wildcard_quantifier_pattern(ast: Node) -> bool:
    """dfa.mojo:3386-3415."""
    if ast.type != RE or ast.get_children_len() != 1:
        return False
    group = ast.get_child(0)
    if group.type != GROUP or group.get_children_len() != 1:
        return False
    w = group.get_child(0)
    if w.type != WILDCARD:
        return False
    return (w.min, w.max) in ((0, -1), (1, -1), (0, 1), (1, 1))


def _inner_or_of_double_group(ast: Node) -> Optional[Node]:
    if ast.type != RE or ast.get_children_len() != 1:
        return None
    outer = ast.get_child(0)
    if outer.type != GROUP or outer.get_children_len() != 1:
        return None
    inner = outer.get_child(0)
    if inner.type != GROUP or inner.get_children_len() != 1:
        return None
    orn = inner.get_child(0)
    return orn if orn.type == OR else None


def _is_common_prefix_alternation_pattern(ast: Node) -> bool:
    """dfa.mojo:3418-3459."""
    orn = _inner_or_of_double_group(ast)
    if orn is None:
        return False
    branches: List[bytes] = []
    if not _extract_literal_branches_strict(orn, branches):
        return False
    if len(branches) < 2:
        return False
    return len(_common_prefix(branches)) >= 2


def _is_quantified_alternation_group(ast: Node) -> bool:
    """dfa.mojo:3525-3562."""
    orn = _inner_or_of_double_group(ast)
    if orn is None:
        return False
    qg = ast.get_child(0).get_child(0)
    if qg.min == 1 and qg.max == 1:
        return False
    return _extract_literal_branches_strict(orn, [])


def compile_dfa_pattern(ast: Node) -> DFA:
    """dfa.mojo:2385-2496 -- shape dispatcher.  Raises Unsupported where the reference
    raises "Pattern too complex" (-> NFA fallback, matcher.mojo:665-675)."""
    dfa = DFA()
    if is_literal_pattern(ast):
        hs, he = pattern_has_anchors(ast)
        compile_pattern(dfa, get_literal_string(ast), hs, he)
    elif _is_pure_anchor_pattern(ast):
        hs, he = pattern_has_anchors(ast)
        compile_pattern(dfa, b"", hs, he)
    elif _is_simple_character_class_pattern(ast):
        # _extract_character_class_info, dfa.mojo:2532-2602
        hs = he = False
        if ast.type in (DIGIT, WORD, RANGE):
            cn = ast
        elif ast.type == RE and ast.get_children_len() == 1:
            c0 = ast.get_child(0)
            if c0.type in (DIGIT, WORD, RANGE):
                cn = c0
            elif c0.type == GROUP and c0.get_children_len() == 1:
                cn = c0.get_child(0)
            else:
                cn = c0
            hs, he = pattern_has_anchors(ast)
        else:
            cn = ast
        if cn.type == DIGIT:
            cc = DIGITS
        elif cn.type == WORD:
            cc = WORD_CHARS
        else:
            cc = cn.get_value() or b""
        compile_character_class_with_logic(dfa, expand_character_range(cc), cn.min, cn.max,
                                           cn.positive_logic)
        dfa.has_start_anchor, dfa.has_end_anchor = hs, he
    elif _is_multi_character_class_sequence(ast):
        els, hs, he = _extract_multi_class_sequence_info(ast)
        compile_multi_character_class_sequence(dfa, els, hs, he)
    elif _is_sequential_character_class_pattern(ast):
        # _extract_sequential_pattern_info, dfa.mojo:2666-2708
        hs, he = pattern_has_anchors(ast)
        els = []
        for e in ast.get_child(0).children:
            if e.type == DIGIT:
                cc = DIGITS
            elif e.type == WORD:
                cc = WORD_CHARS
            elif e.type == RANGE:
                cc = expand_character_range(e.get_value() or b"")
            else:
                continue
            els.append(SeqElement(cc, e.min, e.max, e.positive_logic))
        compile_sequential_pattern(dfa, els, hs, he)
    elif _is_mixed_sequential_pattern(ast):
        els, hs, he = _extract_multi_class_sequence_info(ast)
        compile_multi_character_class_sequence(dfa, els, hs, he)
    elif _is_alternation_pattern(ast):
        compile_alternation(dfa, ast)
        dfa.has_start_anchor, dfa.has_end_anchor = pattern_has_anchors(ast)
    elif _is_quantified_group(ast):
        compile_quantified_group(dfa, ast)
        dfa.has_start_anchor, dfa.has_end_anchor = pattern_has_anchors(ast)
    elif _is_simple_quantifier_pattern(ast):
        compile_simple_quantifier(dfa, ast)
        dfa.has_start_anchor, dfa.has_end_anchor = pattern_has_anchors(ast)
    elif _is_wildcard_quantifier_pattern(ast):
        compile_wildcard_quantifier(dfa, ast)
        dfa.has_start_anchor, dfa.has_end_anchor = pattern_has_anchors(ast)
    elif _is_common_prefix_alternation_pattern(ast):
        compile_common_prefix_alternation(dfa, ast)
        dfa.has_start_anchor, dfa.has_end_anchor = pattern_has_anchors(ast)
    elif _is_quantified_alternation_group(ast):
        compile_quantified_alternation_group(dfa, ast)
        dfa.has_start_anchor, dfa.has_end_anchor = pattern_has_anchors(ast)
    else:
        raise Unsupported("Pattern too complex for current DFA implementation")
    return dfa


# --------------------------------------------------------------------------- #
# HybridMatcher / CompiledRegex construction: matcher.mojo:566-693, 964-1035
def find_rare_required_byte(ast: Node, first_class_lookup: List[int]) -> int:
    """matcher.mojo:122-165."""
    node = ast
    if node.type == RE and node.get_children_len() == 1:
        node = node.get_child(0)
    if node.type != GROUP:
        return -1
    for child in node.children:
        if child.type != ELEMENT or child.min < 1:
            continue
        v = child.get_value()
        if not v or len(v) != 1:
            continue
        if first_class_lookup[v[0]] == 0:
            return v[0]
    return -1


def detect_fixed_width_groups(pattern: bytes) -> Optional[List[int]]:
    """matcher.mojo:1486-1586."""
    p, n = pattern, len(pattern)
    segs: List[int] = []
    i = 0
    lit = 0
    while i < n:
        if p[i] == ord("("):
            if lit > 0:
                segs.append(-lit)
                lit = 0
            if i + 1 < n and p[i + 1] == ord("?"):
                return None
            i += 1
            if i + 1 >= n or p[i] != ord("\\") or p[i + 1] != ord("d"):
                return None
            i += 2
            if i < n and p[i] == ord("{"):
                i += 1
                ns = i
                while i < n and ord("0") <= p[i] <= ord("9"):
                    i += 1
                if i == ns or i >= n or p[i] != ord("}"):
                    return None
                segs.append(int(p[ns:i]))
                i += 1
            elif i < n and p[i] == ord(")"):
                segs.append(1)
            else:
                return None
            if i >= n or p[i] != ord(")"):
                return None
            i += 1
        elif p[i] in (ord("|"), ord("[")):
            return None
        else:
            if p[i] == ord("\\") and i + 1 < n:
                lit += 1
                i += 2
            else:
                lit += 1
                i += 1
    if not any(s > 0 for s in segs):
        return None
    return segs


@dataclass
class Compiled:
    """What CompiledRegex + HybridMatcher hold after construction, restricted to the
    DFA-routed world (matcher.mojo:535-565, 929-1035)."""
    pattern: bytes
    is_wildcard_match_any: bool = False
    complexity: int = SIMPLE
    use_pure_dfa: bool = False
    dfa: Optional[DFA] = None
    prefilter_literal: Optional[bytes] = None  # MemchrPrefilter literal (>= 2 bytes)
    has_anchors_info: bool = False  # literal_info.has_anchors (only set when analysed)
    is_exact_literal: bool = False
    exact_literal: bytes = b""
    required_byte: int = -1
    fixed_sub_total_width: int = -1
    fixed_sub_num_groups: int = 0
    fixed_sub_offsets: List[int] = field(default_factory=lambda: [0] * 10)
    fixed_sub_widths: List[int] = field(default_factory=lambda: [0] * 10)
    fixed_sub_concat: bool = False
    nfa_routed: bool = False      # served by the reference's LazyDFA (oracle/nfa.py); `dfa` then holds its state sets
    program: object = None        # ... and this the PikeVM program


def _route_to_lazy_dfa(c: Compiled, ast: Node, why: str) -> None:
    """matcher.mojo:665-675: no DFAEngine -> NFAMatcher.  Served here when that means the LazyDFA (oracle/nfa.py)."""
    import sys

    from . import nfa

    prog, reason = nfa.route(c.pattern, ast, sys.modules[__name__])
    if prog is None:
        raise Unsupported("%s; %s" % (why, reason))
    built = nfa.subset_construction(prog)
    if built is None:
        raise Unsupported("%s; more than 254 LazyDFA states" % why)
    d = DFA()
    d.trans, d.accepting, d.start_state = built
    c.dfa = d
    c.nfa_routed = True
    c.program = prog


def compile_regex(pattern: bytes) -> Compiled:
    """matcher.mojo:566-693 (HybridMatcher.__init__) + :964-1035 (CompiledRegex)."""
    c = Compiled(pattern=pattern)
    if pattern == b".*":  # matcher.mojo:573-591
        c.is_wildcard_match_any = True
        _precompute_fixed_sub(c)
        return c
    ast = parse(pattern)
    c.complexity = classify(ast)
    c.use_pure_dfa = should_use_pure_dfa(ast)
    analyse = (not is_simple_pattern_skip_prefilter(pattern)) and not c.use_pure_dfa
    if analyse:  # matcher.mojo:609-654
        has_anchors = check_ast_for_anchors(ast)
        best = best_literal(ast)
        best_opt: Optional[bytes] = None
        is_exact = False
        if best is not None and len(best.literal) > 0:
            best_opt = best.literal
            has_ops = any(ch in pattern for ch in b"*+?.|([{")
            is_exact = (best.is_required and has_literal_prefix(ast) and not has_anchors
                        and not has_ops)
        c.has_anchors_info = has_anchors
        c.is_exact_literal = is_exact
        c.exact_literal = best_opt or b""
        if b"|" not in pattern and best_opt is not None and len(best_opt) >= 2:
            c.prefilter_literal = best_opt
    if c.complexity != SIMPLE:
        _route_to_lazy_dfa(c, ast, "pattern classified %s -> NFA engines" % ("MEDIUM" if c.complexity == MEDIUM else "COMPLEX"))
    else:
        try:
            c.dfa = compile_dfa_pattern(ast)
        except Unsupported as e:  # DFA compilation raised: NFA only (matcher.mojo:668-672)
            _route_to_lazy_dfa(c, ast, str(e))
    if (not c.nfa_routed) and (not c.has_anchors_info) and c.dfa.matcher is not None:  # matcher.mojo:684-693
        c.required_byte = find_rare_required_byte(ast, c.dfa.matcher.lookup)
    _precompute_fixed_sub(c)
    return c


def _precompute_fixed_sub(c: Compiled) -> None:
    """matcher.mojo:1002-1035."""
    segs = detect_fixed_width_groups(c.pattern)
    if segs is None:
        return
    ng = tw = 0
    has_lit = False
    offsets = list(c.fixed_sub_offsets)
    widths = list(c.fixed_sub_widths)
    for s in segs:
        if s > 0:
            ng += 1
            if ng > 9:
                # the reference has already written offsets/widths for groups 1..9 but
                # returns before publishing num_groups/total_width (matcher.mojo:1022-1023)
                return
            offsets[ng] = tw
            widths[ng] = s
            tw += s
        else:
            has_lit = True
            tw += -s
    if ng == 0:
        return
    c.fixed_sub_offsets, c.fixed_sub_widths = offsets, widths
    c.fixed_sub_num_groups = ng
    c.fixed_sub_total_width = tw
    c.fixed_sub_concat = not has_lit


# --------------------------------------------------------------------------- #
# Replacement template: matcher.mojo:1437-1482
def has_group_refs(repl: bytes) -> bool:
    """matcher.mojo:1473-1482."""
    for i in range(len(repl) - 1):
        if repl[i] == ord("\\") and ord("1") <= repl[i + 1] <= ord("9"):
            return True
    return False


def parse_repl_template(repl: bytes) -> List[Tuple[int, int, int]]:
    """matcher.mojo:1437-1469 -> list of (group_ref, start, length)."""
    segs: List[Tuple[int, int, int]] = []
    i = 0
    lit_start = 0
    n = len(repl)
    while i < n:
        if repl[i] == ord("\\") and i + 1 < n:
            nx = repl[i + 1]
            if ord("1") <= nx <= ord("9"):
                if i > lit_start:
                    segs.append((0, lit_start, i - lit_start))
                segs.append((nx - ord("0"), 0, 0))
                i += 2
                lit_start = i
                continue
        i += 1
    if lit_start < n:
        segs.append((0, lit_start, n - lit_start))
    return segs
