#!/usr/bin/env python3
"""profiles/r02_traffic.json from `ncu --set full` captures: DRAM bytes per launch of each config's dominant kernel.
usage: make_traffic.py c2=gpurun_out/x.ncu-rep:1073741824 c4=...:4294967296 ...   (config=report:bytes_per_gpu)
Every entry carries bench.kernel_source_hash() of the tree it was made in: bench.py quotes an entry only while the
kernel sources are the ones that were profiled."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
LABEL = {"k_sand": "k_sand", "k_split": "k_split", "k_subst": "k_subst", "k_lines_req": "k_lines_req", "k_runs": "k_runs"}


def main():
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        config, rest = arg.split("=")
        rep, nbytes = rest.rsplit(":", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, row = rows[0], rows[1], rows[2]
        get = lambda k: float(row[hdr.index(k)]) * UNIT[units[hdr.index(k)]]
        name = row[hdr.index("Kernel Name")]
        kernel = next(v for k, v in LABEL.items() if k in name)
        rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        out[config] = {"kernel": kernel, "bytes_per_gpu": int(nbytes), "traffic_bytes": rd + wr, "dram_read": rd,
                       "dram_write": wr, "kernel_name": name, "capture": os.path.basename(rep), "src_sha": kernel_source_hash(kernel)}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
