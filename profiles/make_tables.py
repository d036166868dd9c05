#!/usr/bin/env python3
"""Rewrite the result tables of BASELINE.md (section 3) and README.md from profiles/r02_bench_n{1,2,4,8}.json and
profiles/r02_bench_ref_n1.json (between the <!-- r02-table --> markers).  usage: python profiles/make_tables.py"""
import json, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"c1": "C1 findall `\\d+`, 64 MiB", "c2": "C2 findall `(hello\\|world\\|foo)`, 1 GiB", "c3": "C3 search e-mail, 10 M × 128 B",
         "c4": "C4 sub phone rewrite, 4 GiB", "c5": "C5 split `\\s*,\\s*`, 4 GiB per GPU (32 GiB at N = 8)"}
ALG = {"c1": "N + 16·M", "c2": "N + 16·M", "c3": "N + 16·L", "c4": "N + N_out", "c5": "N + Σpieces + 8·(P+1)"}


def load(n):
    p = os.path.join(ROOT, "profiles", "r02_bench_n%d.json" % n)
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    out = {d["config"]["config"]: d}
    out.update(d.get("configs", {}))
    # k_lines_req got its ticket counter after the full N > 1 lines were taken: C3 at N > 1 comes from a later run of
    # the same command restricted to --configs c3 (r02_bench_n{N}_c3.json); the older value stays in the full line
    o = os.path.join(ROOT, "profiles", "r02_bench_n%d_c3.json" % n)
    if os.path.exists(o):
        out["c3"] = json.load(open(o))["configs"]["c3"]
    return out


def main():
    runs = {n: load(n) for n in (1, 2, 4, 8)}
    ref = None
    p = os.path.join(ROOT, "profiles", "r02_bench_ref_n1.json")
    if os.path.exists(p):
        ref = json.load(open(p))
    lines = ["| config | GPUs | input bytes per GPU | ms per step (device, max over ranks) | input GB/s (whole job) | scaling vs N = 1 | algorithmic bytes | dominant kernel, ms | % of HBM roofline (6467.7 GB/s measured per GPU; in brackets: of 8000 nominal) | DRAM traffic ÷ algorithmic bytes (ncu) | e2e GB/s (host buffers, whole job) | CPU port GB/s (1 thread) |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for c in ("c1", "c2", "c3", "c4", "c5"):
        base = runs[1][c]["value"] if runs[1] and c in runs[1] else None
        for n in (1, 2, 4, 8):
            if not runs[n] or c not in runs[n]:
                continue
            v = runs[n][c]
            r = v["roofline"]
            traffic = "%.3f" % (r["traffic"] / r["algorithmic_bytes_per_launch"]) if r.get("traffic") else "–"
            cpu = v.get("cpu_baseline") or {}
            lines.append("| %s | %d | %d | %.4f | %.0f | %s | %s | `%s` %.4f | %.1f %% (%.1f %%) | %s | %.1f | %s |" % (
                NAMES[c] if n == 1 else "", n, v["config"]["bytes_per_gpu"], v["ms_per_step"], v["value"],
                "%.0f %%" % (100.0 * v["value"] / (n * base)) if base else "–", ALG[c] if n == 1 else "", r["kernel"], r["kernel_ms"],
                100.0 * r["frac"], 100.0 * r["achieved"] / 8000.0, traffic, v["e2e"]["value"], "%.3f" % cpu["value"] if cpu.get("value") else "–"))
    if ref:
        lines.append("")
        lines.append("Reference arm (`bench.py --impl reference`, the oracle port on all %d host threads of the GPU box, %s): %.2f GB/s."
                     % (ref["cpu_baseline"]["cores"], ref["config"]["config"].upper(), ref["value"]))
    table = "\n".join(lines)
    for name in ("BASELINE.md", "README.md"):
        path = os.path.join(ROOT, name)
        s = open(path).read()
        s2 = re.sub(r"<!-- r02-table:begin -->.*?<!-- r02-table:end -->", "<!-- r02-table:begin -->\n" + table.replace("\\", "\\\\") + "\n<!-- r02-table:end -->", s, flags=re.S)
        open(path, "w").write(s2)
    print(table)


if __name__ == "__main__":
    main()
