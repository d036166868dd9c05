#!/usr/bin/env python3
"""Per CUDA source line: warp-instructions executed and stall samples (ncu --page source --print-source cuda,sass).
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; ins = r.index("Instructions Executed"); sam = r.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try:
            key = (cur_file, int(r[0]))
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += int(r[ins] or 0); a[1] += int(r[sam] or 0)
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100.0 * a[0] / max(tot_i, 1), 100.0 * a[1] / max(tot_s, 1), f, ln, a[2]))
