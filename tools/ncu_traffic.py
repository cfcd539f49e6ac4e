#!/usr/bin/env python
"""profiles/r<N>_traffic.json from the .ncu-rep files of one measurement session (tools/gpu_session.sh <tag> ncu): per workload and
kernel, the DRAM bytes of one launch (dram__bytes_read.sum + dram__bytes_write.sum), the frames that launch covered and the counters
that name its limiter.  bench.py reads the newest profiles/r*_traffic.json for `roofline.traffic` / `roofline.limiter`, so every
number in a bench line traces to a tracked file produced by this script from captures of the same build.

Usage: python tools/ncu_traffic.py <tag> <out.json>        (reads gpurun_out/<tag>_*.ncu-rep, writes <out.json> and one
       profiles/<tag>_ncu_<name>_summary.txt per capture)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# capture name (tools/gpu_session.sh) -> (workload, frames per launch of the captured kernel)
# (tools/quick_ab.py renders the spin workload in 120-frame device-pointer calls = one launch sequence of 120 frames; bench.py's 720-frame calls run 240 per sequence)
CAPTURES = {"shade": ("spin1080p", 120), "raster": ("spin1080p", 120), "vertex": ("spin1080p", 120), "setup_spin": ("spin1080p", 120), "prepare": ("spin1080p", 120),
            "setup": ("tess4k", 1), "shade_tess": ("tess4k", 1), "setup50m": ("tess4k_64lights", 1), "raster_over": ("overdraw8k", 1), "shade_over": ("overdraw8k", 1)}
STALLS = ['long_scoreboard', 'wait', 'short_scoreboard', 'branch_resolving', 'no_instruction', 'barrier', 'not_selected', 'lg_throttle',
          'math_pipe_throttle', 'dispatch_stall', 'mio_throttle', 'drain', 'membar', 'imc_miss', 'tex_throttle', 'sleeping']


def rows_of(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(r, units))) for r in rows[2:]]


def scaled(v, unit, kind):
    x = float(v.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12} if kind == "bytes" else {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
    return x * mult.get(unit, 1)


def main():
    tag, out = sys.argv[1], sys.argv[2]
    result = {"source": "gpurun_out/%s_*.ncu-rep via tools/ncu_traffic.py (ncu --set full --clock-control none; one launch each); summaries: profiles/%s_ncu_*_summary.txt" % (tag, tag)}
    for name, (workload, frames) in CAPTURES.items():
        rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, name))
        if not os.path.exists(rep):
            continue
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "profiles", "%s_ncu_%s_summary.txt" % (tag, name)), "w").write(summ)
        for r in rows_of(rep):
            kname = r["Kernel Name"][0].split("(")[0].replace("void ", "").replace("rk::", "")
            kname = kname.split("<")[0]
            g = lambda k, kind=None: (scaled(*r[k], kind) if kind else float(r[k][0].replace(",", ""))) if k in r else None
            stalls = sorted(((float(r['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s][0]), s) for s in STALLS
                             if 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s in r), reverse=True)
            result.setdefault(workload, {})[kname] = {
                "dram_bytes_per_launch": int(g("dram__bytes_read.sum", "bytes") + g("dram__bytes_write.sum", "bytes")),
                "dram_read_bytes": int(g("dram__bytes_read.sum", "bytes")), "dram_write_bytes": int(g("dram__bytes_write.sum", "bytes")),
                "frames_per_launch": frames, "duration_ms_under_ncu": g("gpu__time_duration.sum", "time"),
                "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "l1tex_throughput_pct": g("l1tex__throughput.avg.pct_of_peak_sustained_active"),
                "warp_instructions_per_launch": int(g("smsp__inst_executed.sum")), "lanes_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
                "registers": int(g("launch__registers_per_thread")), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "top_stall": "%s %.2f" % (stalls[0][1], stalls[0][0]) if stalls else None, "capture": "%s_%s.ncu-rep" % (tag, name)}
    json.dump(result, open(out, "w"), indent=1)
    print(json.dumps(result, indent=1))


if __name__ == "__main__":
    main()
