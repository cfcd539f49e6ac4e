"""Golden hashes of BASELINE.json's full-size synthetic workloads (configs 3, 4 and 5), produced by the REFERENCE's own
compiled hot path (oracle/_ref/libref.so) and cross-checked against the oracle on the same inputs.  Runs only in
the dev container (needs /root/reference for libref.so); takes a few minutes and ~12 GB of memory.

  large_cases.json   per case: image size, triangles, FNV-1a-64 of frame (planar RGB8), depth (f32) and of the
                     oracle's winning-triangle ids, visible pixel count

Usage:  python tests/golden/make_golden_large.py [case ...]
The GPU parity tests (tests/test_parity_gpu_large.py) regenerate the same inputs with rasteriser_b200/synth.py and
compare hashes, so the full-size frames are pinned without running a CPU renderer for minutes on the GPU box."""
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
import scenes as S  # noqa: E402
from rasteriser_b200 import synth  # noqa: E402
from test_oracle_vs_reference import _ref_draw  # noqa: E402


def cases():
    base = S.scene("suzanne")
    yield ("config3_tess91_4k", lambda: orc.Scene(*synth.tessellate(base.positions, base.normals, base.uvs, base.tris, 91), base.materials),
           S.lights("threepoint"), orc.make_args(3840, 2160), 1)
    yield ("config4_overdraw_8k", lambda: orc.Scene(*synth.overdraw_scene(200000, 7680, 4320), [{"kd": (0.8, 0.8, 0.8), "texels": None}]),
           S.lights("threepoint"), orc.make_args(7680, 4320), 8)
    yield ("config5_tess227_64lights_4k", lambda: orc.Scene(*synth.tessellate(base.positions, base.normals, base.uvs, base.tris, 227), base.materials),
           synth.random_lights(64), orc.make_args(3840, 2160), 1)


def main():
    path = os.path.join(HERE, "large_cases.json")
    only = sys.argv[1:]  # case names to (re)generate; default: all
    out = json.load(open(path)) if only and os.path.exists(path) else {}
    assert orc.ref() is not None, "oracle/_ref/libref.so is needed (python -c 'import __graft_entry__ as g; g.build()')"
    for name, make, lights, args, threads in cases():
        if only and name not in only:
            continue
        t0 = time.time()
        scene = make()
        print(name, "scene", len(scene.tris), "triangles, %.0f s" % (time.time() - t0), flush=True)
        t0 = time.time()
        f, d, t = orc.oracle_draw(scene, lights, args, threads=threads)
        print("  oracle %.0f s" % (time.time() - t0), flush=True)
        t0 = time.time()
        with tempfile.TemporaryDirectory() as tmp:
            rf, rd = _ref_draw(scene, lights, args, tmp)
        print("  reference %.0f s" % (time.time() - t0), flush=True)
        assert np.array_equal(f, rf) and np.array_equal(d.view(np.uint32), rd.view(np.uint32)), name + ": oracle != reference"
        out[name] = {"width": int(args.image_width), "height": int(args.image_height), "triangles": int(len(scene.tris)),
                     "frame_fnv": orc.fnv(rf), "depth_fnv": orc.fnv(rd), "tri_fnv": orc.fnv(t), "visible_pixels": int((t != orc.NO_TRIANGLE).sum()),
                     "produced_by": "reference (oracle/_ref/libref.so); oracle identical"}
        del scene, f, d, t, rf, rd
        json.dump(out, open(os.path.join(HERE, "large_cases.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
