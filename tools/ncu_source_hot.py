#!/usr/bin/env python
"""Where a kernel's instructions and stall samples go, from the source page of an .ncu-rep (ncu --set full --import-source on):
the SASS split into runs of equal execution count, with warp instructions executed, average active lanes and stall samples per run.
Usage: python tools/ncu_source_hot.py file.ncu-rep [--all]"""
import csv
import io
import subprocess
import sys

txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:150])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ins.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), float(r[ix["Avg. Threads Executed"]] or 0), int(r[ix["# Samples"]])))
tot_i = sum(i[1] for i in ins)
tot_s = sum(i[3] for i in ins)
print("instructions %d, warp instructions executed %d, samples %d" % (len(ins), tot_i, tot_s))
if "--all" in sys.argv:
    for k, (s, n, t, sm) in enumerate(ins):
        print("%4d %-60s %10d %5.1f %6d" % (k, s[:60], n, t, sm))
    sys.exit()
run = []
def flush():
    if not run:
        return
    n = sum(i[1] for _, i in run)
    s = sum(i[3] for _, i in run)
    lanes = sum(i[1] * i[2] for _, i in run) / max(n, 1)
    ops = {}
    for _, i in run:
        o = i[0].split()[0] if not i[0].startswith("@") else i[0].split()[1]
        o = o.split(".")[0]
        ops[o] = ops.get(o, 0) + 1
    top = " ".join("%s%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print("%4d-%4d  n=%3d  exec/instr %9d  inst %5.1f%%  lanes %4.1f  samples %5.1f%%   %s" % (run[0][0], run[-1][0], len(run), run[0][1][1], 100.0 * n / tot_i, lanes, 100.0 * s / max(tot_s, 1), top))
for k, i in enumerate(ins):
    if run and abs(i[1] - run[0][1][1]) > 0.02 * max(run[0][1][1], 1):
        flush()
        run = []
    run.append((k, i))
flush()
