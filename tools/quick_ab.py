#!/usr/bin/env python
"""Torch-free timing + output hash of one workload through the C ABI (starts in about a second: for short GPU sessions and
A/B runs of kernel variants -- RAST_LIB=build/variants/librast_b200_<name>.so python tools/quick_ab.py spin1080p).

Renders `frames` poses per call into device memory (rast_device_alloc), `--calls` timed calls after 3 warm-up calls:
wall clock around the calls + rast_sync (two-stream overlap on, the way bench.py's `value` runs), then the same with
rast_set_profiling for per-pass device times.  The hash (FNV-1a of the first frames' RGB planes, their depth planes and
the last frame's triangle ids, read back with rast_device_read) must be equal between variants: same inputs, same bits.
Prints one JSON line.  Not a bench value (no clocks sampling, no barrier protocol): bench.py is the benchmark."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def fnv1a(a):
    import orc
    return orc.fnv(a)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["spin1080p", "suzanne640", "tess4k", "tess4k_64lights", "overdraw8k"])
    ap.add_argument("--calls", type=int, default=5)
    ap.add_argument("--frames", type=int, default=0, help="frames per call (default: 120 for the spin sequence, 1 otherwise)")
    opts = ap.parse_args()
    import bench
    from rasteriser_b200 import api
    wl = bench.make_workload(opts.workload)
    n = opts.frames or (120 if wl["frames"] > 1 else 1)
    wl["frames"] = n
    W, H = wl["width"], wl["height"]
    P = W * H
    r = api.Renderer(0)
    r.upload_mesh(wl["pos"], wl["tris"], wl["nrm"], wl["uv"])
    r.upload_materials(wl["materials"])
    r.set_lights(wl["lights"])
    poses = bench.spin_args(api, wl, 0, 1)
    arr = (api.RastArgs * n)(*[a.to_rast() for a in poses])
    frames_dev, depths_dev = r.device_alloc(n * 3 * P), r.device_alloc(n * 4 * P)

    def call():
        r.draw_frames_device(arr, frames_dev, depths_dev)

    for _ in range(3):
        call()
        r.sync()
    t0 = time.perf_counter()
    for _ in range(opts.calls):
        call()
    r.sync()
    wall_ms = (time.perf_counter() - t0) * 1e3 / opts.calls
    r.set_profiling(True)
    passes = {}
    for _ in range(opts.calls):
        call()
        r.sync()
        for k, v in r.pass_ms().items():
            passes[k] = passes.get(k, 0.0) + v / opts.calls
    r.set_profiling(False)
    r.set_keep_visibility(True)
    call()
    r.sync()
    k = min(n, 4)
    rgb, dep = np.empty((k, 3, H, W), np.uint8), np.empty((k, H, W), np.float32)
    r.device_read(frames_dev, rgb)
    r.device_read(depths_dev, dep)
    ids = r.triangle_ids(W, H)
    out = {"workload": opts.workload, "lib": os.path.basename(os.environ.get("RAST_LIB", "librast_b200.so")), "frames_per_call": n, "calls": opts.calls,
           "ms_per_call": round(wall_ms, 4), "frames_per_s": round(n / wall_ms * 1e3, 1), "pass_ms_per_call": {k_: round(v, 4) for k_, v in passes.items()},
           "hash_rgb": fnv1a(rgb), "hash_depth": fnv1a(dep), "hash_ids": fnv1a(ids), "visible_last": int((ids != 0xFFFFFFFF).sum()), "stats": r.stats()}
    print(json.dumps(out), flush=True)
    r.device_free(frames_dev)
    r.device_free(depths_dev)
    r.close()


if __name__ == "__main__":
    main()
