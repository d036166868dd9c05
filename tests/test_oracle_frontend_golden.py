"""Pin the oracle's front end (oracle/frontend.py: lexer, parser, classifier, literal helpers)
to the assertions of the reference's own front-end tests.

Hand transcriptions of /root/reference/tests/test_lexer.mojo, test_parser.mojo and
test_optimizer.mojo (file:line beside each case).  The reference cannot run here (Mojo), so
these fixtures are what ties the restated front end -- which decides the AST shape, hence the
automaton, hence whether a pattern is DFA-routed at all -- to the reference.  The product's
C++ front end is tied to this one by tests/test_cabi_cpu.py (table-for-table on 420 random
patterns plus every golden pattern).
"""
import pytest

from oracle import frontend as fe

TL = "tests/test_lexer.mojo"
TPA = "tests/test_parser.mojo"
TO = "tests/test_optimizer.mojo"


def top(pattern):
    return fe.parse(pattern.encode()).get_child(0)


def in_class(node, ch):
    """Membership the DFA builders derive for a RANGE node: dfa.mojo:71-168 expansion,
    inverted when the node carries negative logic (dfa.mojo:1747-1803)."""
    members = fe.expand_character_range(node.get_value())
    return (ord(ch) in members) == node.positive_logic


# ---- lexer ---------------------------------------------------------------------------------
LEX_TYPES = [
    (f"{TL}:29-32", "a", {0: fe.T_ELEMENT}),
    (f"{TL}:35-38", "a\\\\a\\\\t\\.", {1: fe.T_ELEMENT}),
    (f"{TL}:41-44", "a\\h\\t", {2: fe.T_ELEMENT}),
    (f"{TL}:47-50", "\\.", {0: fe.T_ELEMENT}),
    (f"{TL}:53-55", "a{3,5}", {3: fe.T_COMMA}),
    (f"{TL}:58-60", "a,", {1: fe.T_ELEMENT}),
    (f"{TL}:63-65", "^a", {0: fe.T_START}),
    (f"{TL}:68-70", "fdsad\\$cs$", {-1: fe.T_END}),
    (f"{TL}:78-85", "-\\\\/\\s~", {0: fe.T_DASH, 1: fe.T_ELEMENT, 2: fe.T_ELEMENT, 3: fe.T_SPACE, 4: fe.T_ELEMENT}),
    (f"{TL}:88-90", ".", {0: fe.T_WILDCARD}),
    (f"{TL}:94-97", "()", {0: fe.T_LPAREN, 1: fe.T_RPAREN}),
    (f"{TL}:100-103", "[]", {0: fe.T_LBRACKET, 1: fe.T_RBRACKET}),
    (f"{TL}:106-110", "{1}", {0: fe.T_LCURLY, 1: fe.T_ELEMENT, 2: fe.T_RCURLY}),
    (f"{TL}:113-117", "*+?", {0: fe.T_ASTERISK, 1: fe.T_PLUS, 2: fe.T_QMARK}),
    (f"{TL}:120-122", "|", {0: fe.T_VBAR}),
    (f"{TL}:125-127", "-", {0: fe.T_DASH}),
]


@pytest.mark.parametrize("src,pattern,types", LEX_TYPES, ids=[c[0] for c in LEX_TYPES])
def test_lexer_token_types(src, pattern, types):
    tokens = fe.scan(pattern.encode())
    for idx, ty in types.items():
        assert tokens[idx].type == ty, (src, pattern, idx)


def test_lexer_token_chars_and_counts():
    assert fe.scan(b"a")[0].char == ord("a")                       # test_lexer.mojo:31
    assert fe.scan(b"a\\\\a\\\\t\\.")[1].char == ord("\\")          # :38
    assert fe.scan(b"a\\h\\t")[2].char == 9                        # :44
    assert fe.scan(b"\\.")[0].char == ord(".")                     # :50
    assert len(fe.scan(b"-\\\\/\\s~")) == 5                        # :80


def test_lexer_bad_curly_raises():
    with pytest.raises(fe.RegexError):                              # test_lexer.mojo:73-75
        fe.scan(b"advfe{a}")


# ---- parser --------------------------------------------------------------------------------
def test_simple_regex():                                            # test_parser.mojo:23-33
    ast = fe.parse(b"a")
    assert ast.type == fe.RE and ast.get_children_len() == 1
    child = ast.get_child(0)
    assert child.type == fe.GROUP and child.get_children_len() == 1
    element = child.get_child(0)
    assert element.type == fe.ELEMENT and element.get_value() == b"a"


def test_grouping():                                                # test_parser.mojo:36-51
    g = top("a(b)c")
    assert g.get_children_len() == 3
    assert [c.type for c in g.children] == [fe.ELEMENT, fe.GROUP, fe.ELEMENT]
    nested = g.get_child(1)
    assert nested.get_children_len() == 1
    assert nested.get_child(0).type == fe.ELEMENT
    assert nested.get_child(0).get_value() == b"b"


QUANT = [
    (f"{TPA}:54-60", "a*", fe.ELEMENT, 0, -1),
    (f"{TPA}:110-114", "a{3}", fe.ELEMENT, 3, 3),
    (f"{TPA}:117-122", "a{2,5}", fe.ELEMENT, 2, 5),
    (f"{TPA}:138-144", "a{,2}", fe.ELEMENT, 0, 2),
    (f"{TPA}:147-153", "a{2,}", fe.ELEMENT, 2, -1),
    (f"{TPA}:156-162", "a{,}", fe.ELEMENT, 0, -1),
]


@pytest.mark.parametrize("src,pattern,ty,mn,mx", QUANT, ids=[c[1] for c in QUANT])
def test_quantifier_bounds(src, pattern, ty, mn, mx):
    e = top(pattern).get_child(0)
    assert (e.type, e.min, e.max) == (ty, mn, mx), src


def test_match_start_end():                                         # test_parser.mojo:63-69
    g = top("^a$")
    assert [c.type for c in g.children] == [fe.START, fe.ELEMENT, fe.END]


def test_wildcard_and_space_nodes():                                # test_parser.mojo:72-83
    assert top(".").get_child(0).type == fe.WILDCARD
    assert top("\\s").get_child(0).type == fe.SPACE


def test_range_positive():                                          # test_parser.mojo:86-94
    e = top("[abc]").get_child(0)
    assert e.type == fe.RANGE and e.min == 1 and e.positive_logic
    assert in_class(e, "a") and not in_class(e, "x")


def test_range_negative():                                          # test_parser.mojo:97-107
    e = top("[^abc]").get_child(0)
    assert e.type == fe.RANGE and e.min == 1 and not e.positive_logic
    assert not in_class(e, "a") and in_class(e, "x")


def test_or_operation():                                            # test_parser.mojo:110-114
    n = top("a|b")
    assert n.type == fe.OR and n.get_children_len() == 2


def test_curly_braces_1():                                          # test_parser.mojo:131-136
    assert top("a{5}b").get_children_len() == 2


def test_parse_match_start_end():                                   # test_parser.mojo:165-168
    assert top("^aaaa.*a$").get_children_len() == 8


def test_complex_regex():                                           # test_parser.mojo:171-192
    g = top("^[a-zA-Z]{1,20}@[a-zA-Z]\\.[a-z]{1,3}$")
    assert [c.type for c in g.children] == [fe.START, fe.RANGE, fe.ELEMENT, fe.RANGE, fe.ELEMENT,
                                            fe.RANGE, fe.END]
    assert (g.get_child(1).min, g.get_child(1).max) == (1, 20)
    assert (g.get_child(5).min, g.get_child(5).max) == (1, 3)


def test_ranges_with_dashes():                                      # test_parser.mojo:195-224
    e = top("[^a-z]").get_child(0)
    assert e.type == fe.RANGE and e.min == 1 and not in_class(e, "a")
    e = top("[^a-z-\\s-]").get_child(0)
    assert e.type == fe.RANGE and not in_class(e, "a") and not in_class(e, "-")
    e = top("[a-z-\\s-]").get_child(0)
    assert e.type == fe.RANGE and in_class(e, "a") and in_class(e, "-")


@pytest.mark.parametrize("pattern", ["a[abc", "abc]", "a)", "a[d]((vfw)", "[a", "["])
def test_parse_failures(pattern):                                   # test_parser.mojo:227-252
    with pytest.raises(fe.RegexError):
        fe.parse(pattern.encode())


# ---- classifier and literal helpers --------------------------------------------------------
SIMPLE_PATTERNS = [
    (f"{TO}:16-18", "hello"), (f"{TO}:24-25", "^hello"), (f"{TO}:31-33", "hello$"),
    (f"{TO}:41-42", "a*"), (f"{TO}:44-45", "a+"), (f"{TO}:47-48", "a?"), (f"{TO}:50-51", "a{3}"),
    (f"{TO}:53-54", "a{2,5}"), (f"{TO}:60-61", "[a-z]"), (f"{TO}:67-68", "[a-z]+"),
    (f"{TO}:74-75", "[0-9]*"), (f"{TO}:81-82", "[^a-z]"), (f"{TO}:89-90", "(abc)+"),
    (f"{TO}:220", "."), (f"{TO}:226", ".*"), (f"{TO}:232", ".+"), (f"{TO}:238", ".?"),
    (f"{TO}:246-247", "\\s"), (f"{TO}:249-250", "\\s*"), (f"{TO}:252-253", "\\s+"),
]


@pytest.mark.parametrize("src,pattern", SIMPLE_PATTERNS, ids=[c[1] for c in SIMPLE_PATTERNS])
def test_classified_simple(src, pattern):
    assert fe.classify(fe.parse(pattern.encode())) == fe.SIMPLE, src


def test_classified_other():
    assert fe.classify(fe.parse(b"a|b|c")) in (fe.SIMPLE, fe.MEDIUM)          # test_optimizer.mojo:97-103
    assert fe.classify(fe.parse(b"a{50,100}")) in (fe.MEDIUM, fe.COMPLEX)     # :262-268
    assert fe.classify(fe.parse(b"a{1000,2000}")) == fe.COMPLEX               # :270-271


def test_is_literal_pattern():                                      # test_optimizer.mojo:106-156
    for p in ("hello", "abc123", "", "^hello", "hello$", "^hello$"):
        assert fe.is_literal_pattern(fe.parse(p.encode())), p
    for p in ("a*", "a+", "a?", "[a-z]", "a|b", "(abc)"):
        assert not fe.is_literal_pattern(fe.parse(p.encode())), p


def test_get_literal_string():                                      # test_optimizer.mojo:159-186
    for p, lit in (("hello", b"hello"), ("abc123", b"abc123"), ("", b""), ("^hello", b"hello"),
                   ("hello$", b"hello"), ("^hello$", b"hello")):
        assert fe.get_literal_string(fe.parse(p.encode())) == lit, p


def test_pattern_has_anchors():                                     # test_optimizer.mojo:189-214
    assert fe.pattern_has_anchors(fe.parse(b"hello")) == (False, False)
    assert fe.pattern_has_anchors(fe.parse(b"^hello")) == (True, False)
    assert fe.pattern_has_anchors(fe.parse(b"hello$")) == (False, True)
    assert fe.pattern_has_anchors(fe.parse(b"^hello$")) == (True, True)
