#!/usr/bin/env python
"""Output-path benchmark: frame.png of a W x H RGB frame through the product's PNG writer,
one thread against all host threads.  The frame is a rendered-looking synthetic (smooth shading + texture noise on
a black background).  Usage: python tools/bench_output.py [--size 7680x4320].  Prints one JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="7680x4320")
    o = ap.parse_args()
    w, h = [int(v) for v in o.size.split("x")]
    from rasteriser_b200 import hostio
    l = hostio.lib()
    l.rasth_png_set_threads.argtypes = [C.c_uint]
    rng = np.random.RandomState(3)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    inside = ((xx - w / 2) ** 2 / (w * 0.3) ** 2 + (yy - h / 2) ** 2 / (h * 0.4) ** 2) < 1
    shade = np.clip(200 * (1 - ((xx - w * 0.4) ** 2 + (yy - h * 0.4) ** 2) / (w * 0.5) ** 2), 0, 255)
    img = np.stack([shade * 0.5, shade, shade * 0.4]) + rng.randint(0, 12, (3, h, w))
    img = (np.clip(img, 0, 255) * inside[None]).astype(np.uint8)
    out = {"workload": "frame.png of a %dx%d RGB8 frame (planar in, zlib level 1)" % (w, h), "raw_mb": img.size / 1e6, "host_threads": os.cpu_count(), "runs": []}
    with tempfile.TemporaryDirectory() as tmp:
        for threads in (1, 0):
            l.rasth_png_set_threads(threads)
            p = os.path.join(tmp, "f%d.png" % threads)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                assert l.rasth_png_write(p.encode(), img.ctypes.data, w, h, 3) == 0
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out["runs"].append({"threads": threads or os.cpu_count(), "seconds": best, "raw_mb_per_s": img.size / 1e6 / best, "file_mb": os.path.getsize(p) / 1e6})
        l.rasth_png_set_threads(0)
        from PIL import Image
        Image.MAX_IMAGE_PIXELS = None
        with Image.open(os.path.join(tmp, "f0.png")) as im:
            back = np.asarray(im)
        out["decodes_identically"] = bool(np.array_equal(back.transpose(2, 0, 1), img))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
