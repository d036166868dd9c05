"""Pin the oracle (oracle/frontend.py + oracle/scan.c) to the reference's own test vectors."""
import pytest

import oracle
from _golden import check_vector, load_vectors

DATA = load_vectors()


@pytest.mark.parametrize("v", DATA["vectors"], ids=lambda v: "%s:%s" % (v["op"], v["src"].split("/")[-1]))
def test_reference_vector(v):
    check_vector(oracle, v)


@pytest.mark.parametrize("v", DATA["nfa_routed"], ids=lambda v: v["pattern"])
def test_nfa_routed_patterns_are_refused(v):
    with pytest.raises(oracle.Unsupported):
        oracle.compile(v["pattern"])


@pytest.mark.parametrize("v", DATA["group_sub"], ids=lambda v: "%s:%s" % (v["pattern"], v["src"].split("/")[-1]))
def test_group_sub_needs_the_capture_engine(v):
    """sub() with references to groups of no fixed width: the pattern compiles (LazyDFA route), the replacement needs
    NFAEngine.match_next_with_groups (matcher.mojo:1783-1821) -- refused, not guessed."""
    rx = oracle.compile(v["pattern"])
    with pytest.raises(oracle.Unsupported):
        rx.sub(v["repl"].encode(), v["text"].encode())
    assert rx.sub(b"<>", v["text"].encode())  # a replacement without group references is served
