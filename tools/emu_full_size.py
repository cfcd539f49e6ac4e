#!/usr/bin/env python
"""Full-size config 3 (n = 91) / config 5 (n = 227, 64 lights) frames through the host emulation of the kernel functions
(tests/emu_device_fns.cu) with and without the kernel variants, compared with the oracle pixel for pixel.  CPU only.
Usage: python tools/emu_full_size.py [91|227]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import scenes as S  # noqa: E402
import test_emu_device_fns as T  # noqa: E402
from rasteriser_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 91
emu = T.load_emu()
base = S.scene("suzanne")
pos, nrm, uv, tris = synth.tessellate(base.positions, base.normals, base.uvs, base.tris, n)
scene = orc.Scene(pos, nrm, uv, tris, base.materials)
lights = S.lights("threepoint") if n == 91 else synth.random_lights(64)
oa = orc.make_args(3840, 2160)
t = time.time()
want = orc.oracle_draw(scene, lights, oa, threads=1)
print("oracle: %d triangles, %d visible pixels, %.1f s" % (len(tris), int((want[2] != orc.NO_TRIANGLE).sum()), time.time() - t), flush=True)
for name, flags in [("default", 0), ("tight", T.TIGHT), ("tight + prep", T.PRE_NORMALS | T.PREP | T.TIGHT)]:
    t = time.time()
    T.assert_exact(T.emu_draw(emu, scene, lights, oa, flags, 64), want, name)
    print("%s: identical to the oracle (%.1f s)" % (name, time.time() - t), flush=True)
