#!/usr/bin/env python
"""Compare the SASS of every kernel of two builds of librast_b200.so, instruction stream by instruction stream (names may differ in
template arguments: pass name pairs as "new name=>old name").  Used to show that a host-side or structural change left the device code of a
GPU-tested build untouched.  Usage: python tools/sass_compare.py new.so old.so ["new kernel name=>old kernel name" ...]"""
import collections
import re
import subprocess
import sys


def funcs(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out, name = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            out[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and name:
            out[name].append(m.group(2).strip())
    return out


a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
pairs = dict(p.split("=>") for p in sys.argv[3:])
bad = 0
for k, ins in a.items():
    kb = pairs.get(k, k)
    if kb not in b:
        print("only in", sys.argv[1], ":", k, len(ins))
        continue
    same = ins == b[kb]
    bad += not same
    if not same or k in pairs:
        print("%-60s %5d instructions  %s %s" % (k, len(ins), "==" if same else "!=", kb))
print("%d kernels compared, %d differ" % (len(a), bad))
