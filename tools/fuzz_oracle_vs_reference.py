#!/usr/bin/env python
"""CPU-only sweep: the oracle against the reference's own compiled hot path (oracle/_ref/libref.so) over the same
randomised case distribution the GPU fuzz test uses (tests/test_parity_gpu_fuzz.py::_case), so that GPU == oracle
(measured on the B200) and oracle == reference (measured here) cover the same ground.  Dev container only.
Usage: python tools/fuzz_oracle_vs_reference.py [first_seed] [count]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
from test_oracle_vs_reference import _quantised, _ref_draw  # noqa: E402
from test_parity_gpu_fuzz import _case  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
assert orc.ref() is not None, "oracle/_ref/libref.so is needed"
bad = 0
with tempfile.TemporaryDirectory() as tmp:
    for seed in range(first, first + count):
        scene, lights, oa, mode, kind = _case(seed)
        if kind.startswith("soup"):
            scene = _quantised(scene)
        rf, rd = _ref_draw(scene, lights, oa, tmp)
        f, d, t = orc.oracle_draw(scene, lights, oa, threads=2)
        ok = np.array_equal(f, rf) and np.array_equal(d.view(np.uint32), rd.view(np.uint32))
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, kind, oa.image_width, oa.image_height, int((f != rf).sum()), int((d.view(np.uint32) != rd.view(np.uint32)).sum()), flush=True)
print("seeds %d..%d: %d cases, %d mismatches" % (first, first + count - 1, count, bad))
