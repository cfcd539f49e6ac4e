#!/usr/bin/env python
"""Per-kernel SASS statistics of librast_b200.so: instruction count and opcode mix (cuobjdump -sass).
Usage: python tools/sass_stats.py [kernel-substring] [--dump]"""
import collections
import re
import subprocess
import sys

so = __import__("os").environ.get("RAST_LIB") or "rasteriser_b200/librast_b200.so"
pat = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else ""
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, funcs = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        funcs[name + " @" + str(len(funcs))] = []
        cur = funcs[name + " @" + str(len(funcs) - 1)]
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and name:
        cur.append(m.group(2))
for fn, ins in funcs.items():
    if pat not in fn:
        continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in ins)
    print("%s: %d instructions" % (fn, len(ins)))
    print("   " + ", ".join("%s %d" % kv for kv in ops.most_common(18)))
    if "--dump" in sys.argv:
        print("\n".join(ins))
