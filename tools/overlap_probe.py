#!/usr/bin/env python
"""Probe: do two contexts on two streams of one GPU (their raster and shade kernels free to overlap) render the 1080p spin
sequence faster than one context?  Both kernels are issue-bound at 66-76 %; co-residency could fill the gaps."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rasteriser_b200 import api  # noqa: E402

wl = bench.make_workload("spin1080p")
dev = torch.device("cuda", 0)
W, H, n = 1920, 1080, 720


def make(k):
    r = api.Renderer(0)
    r.upload_mesh(wl["pos"], wl["tris"], wl["nrm"], wl["uv"])
    r.upload_materials(wl["materials"])
    r.set_lights(wl["lights"])
    st = torch.cuda.Stream(dev)
    r.set_stream(st.cuda_stream)
    poses = bench.spin_args(api, wl, 0, 1)
    arr = (api.RastArgs * n)(*[a.to_rast() for a in poses])
    f = torch.empty((n, 3, H, W), dtype=torch.uint8, device=dev)
    d = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    return r, st, arr, f, d


ctxs = [make(k) for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2)]
for nctx in range(1, len(ctxs) + 1):
    use = ctxs[:nctx]
    for _ in range(3):
        for r, st, arr, f, d in use:
            r.draw_frames_device(arr, f.data_ptr(), d.data_ptr())
        for r, *_ in use:
            r.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 5
    for _ in range(steps):
        for r, st, arr, f, d in use:
            r.draw_frames_device(arr, f.data_ptr(), d.data_ptr())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%d context(s): %.0f frames/s" % (nctx, nctx * n * steps / dt))
