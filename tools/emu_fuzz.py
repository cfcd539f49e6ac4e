#!/usr/bin/env python
"""The GPU fuzz distribution (tests/test_parity_gpu_fuzz.py::_case) through the host emulation of the kernel functions
(tests/emu_device_fns.cu) with the shipped flags (tight bbox walk, prepared records) and the chunk path with early z,
compared with the oracle bit for bit.  CPU only.  Usage: python tools/emu_fuzz.py [first_seed] [count]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import test_emu_device_fns as T  # noqa: E402
from test_parity_gpu_fuzz import _case  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
emu = T.load_emu()
bad = 0
for seed in range(first, first + count):
    scene, lights, oa, mode, kind = _case(seed)
    want = orc.oracle_draw(scene, lights, oa, threads=2)
    for flags, tiny in [(T.PREP | T.TIGHT, 16), (T.PRE_NORMALS | T.TIGHT, 64), (T.TIGHT | T.ALL_CHUNKS | T.EARLY_Z, 16)]:
        try:
            T.assert_exact(T.emu_draw(emu, scene, lights, oa, flags, tiny), want, "seed %d (%s) flags %d" % (seed, kind, flags))
        except AssertionError as e:
            bad += 1
            print("MISMATCH", e, flush=True)
print("seeds %d..%d: %d cases x 3 flag sets, %d mismatches" % (first, first + count - 1, count, bad))
