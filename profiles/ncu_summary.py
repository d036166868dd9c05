#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + source page) into text: key metrics, stall breakdown,
instruction / sample / shared-wavefront totals per 50-instruction SASS bucket.
usage: ncu_summary.py report.ncu-rep [bucket]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg']
for row in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("%-75s %s %s" % (k, row[i], units[i]))
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and 'per_issue_active' in h and float(row[i] or 0) > 0.05:
            print("%-75s %s" % (h.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''), row[i]))
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for k, r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r:
        h = r; data = rows[k + 1:]; break
if h:
    isrc = h.index('Source'); ins = h.index('Instructions Executed'); isam = h.index('# Samples'); iw = h.index('L1 Wavefronts Shared')
    tot = sum(int(r[ins]) for r in data); ts = sum(int(r[isam]) for r in data)
    print("source page: total warp-instructions %d, samples %d" % (tot, ts))
    print("bucket  instr  samples  smem_wavefronts  first_sass")
    for b in range(0, len(data), B):
        seg = data[b:b + B]
        n = sum(int(r[ins]) for r in seg); s = sum(int(r[isam]) for r in seg); w = sum(int(r[iw] or 0) for r in seg)
        if n > 0.002 * tot or s > 0.005 * ts:
            print(b, n, s, w, seg[0][isrc][:60])
