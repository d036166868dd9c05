"""Shared checker: evaluate the transcribed reference vectors against an implementation.

`impl` is any object with compile(pattern) -> regex exposing
findall(text)->(n,2) int array, search(text)->(s,e)|None, match_first(text)->(s,e)|None,
sub(repl, text, count)->bytes, split(text, maxsplit)->list[bytes].
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_vectors():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)


def check_vector(impl, v):
    rx = impl.compile(v["pattern"])
    text = v["text"].encode()
    op = v["op"]
    if op == "findall":
        spans = [list(map(int, s)) for s in rx.findall(text)]
        if "spans" in v:
            assert spans == v["spans"], (v, spans)
        if "texts" in v:
            assert [text[s:e].decode() for s, e in spans] == v["texts"], (v, spans)
        if "starts" in v:
            assert [s for s, _ in spans] == v["starts"], (v, spans)
        if "count" in v:
            assert len(spans) == v["count"], (v, spans)
        if "last" in v:
            assert spans[-1] == v["last"], (v, spans)
    elif op in ("search", "match_first"):
        r = rx.search(text) if op == "search" else rx.match_first(text)
        if "found" in v:  # CompiledRegex.test / is_match: only whether a match exists is asserted
            assert (r is not None) == v["found"], (v, r)
        elif "starts" in v:
            assert r is not None and r[0] == v["starts"], (v, r)
        elif v["span"] is None:
            assert r is None, (v, r)
        else:
            assert r is not None and list(map(int, r)) == v["span"], (v, r)
    elif op == "sub":
        out = rx.sub(v["repl"].encode(), text, v.get("count", 0))
        assert out == v["out"].encode(), (v, out)
    elif op == "split":
        pieces = rx.split(text, v.get("maxsplit", 0))
        assert [p.decode() for p in pieces] == v["pieces"], (v, pieces)
    else:
        raise AssertionError(op)
