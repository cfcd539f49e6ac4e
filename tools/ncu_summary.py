#!/usr/bin/env python
"""Print the key metrics of every kernel in an .ncu-rep (ncu -i ... --page raw --csv).  Usage: tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'smsp__inst_executed_op_global_red.sum', 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum']
STALLS = ['long_scoreboard', 'wait', 'short_scoreboard', 'branch_resolving', 'no_instruction', 'barrier', 'not_selected', 'lg_throttle',
          'math_pipe_throttle', 'dispatch_stall', 'mio_throttle', 'drain', 'membar', 'imc_miss', 'tex_throttle', 'sleeping']
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("=====", r[hdr.index('Kernel Name')][:110])
    for w in WANT:
        if w in hdr:
            print("  %-60s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = []
    for s in STALLS:
        k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
        if k in hdr:
            st.append((float(r[hdr.index(k)]), s))
    print("  stalls/issue: " + ", ".join("%s %.2f" % (s, v) for v, s in sorted(st, reverse=True)[:7]))
