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
