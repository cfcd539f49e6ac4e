#!/usr/bin/env python
"""Loader benchmark: parse a synthetic tessellated-Suzanne .obj with the product's parallel
reader at several thread counts, with the reference's own loader (oracle/_ref/libref.so = fileloader.cpp +
tinyobjloader, when built) and through the binary mesh cache; checks that all of them give the same arrays.
Usage: python tools/bench_loader.py [--n 32] [--threads 1,2,4,8,0] [--keep DIR]      (n = 91 -> 8.0 M triangles)
Prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--threads", default="1,2,4,8,0")
    ap.add_argument("--keep", default="")
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--cli", action="store_true", help="also time the renderer executable end to end on the scene (needs a GPU)")
    o = ap.parse_args()
    from rasteriser_b200 import hostio, synth
    import orc
    data = os.path.join(ROOT, "tests", "data")
    base, _ = hostio.load_obj(os.path.join(data, "Suzanne.obj"), "")
    pos, nrm, uv, tris = synth.tessellate(base["pos"], base["nrm"], base["uv"], base["tris"], o.n)
    tmp = o.keep or tempfile.mkdtemp()
    os.makedirs(tmp, exist_ok=True)
    path = os.path.join(tmp, "tess%d.obj" % o.n)
    if not os.path.exists(path):
        t0 = time.perf_counter()
        synth.write_obj(path, pos, nrm, uv, tris)
        print("wrote %s (%.1f MB) in %.1f s" % (path, os.path.getsize(path) / 1e6, time.perf_counter() - t0), file=sys.stderr)
    size = os.path.getsize(path)
    out = {"workload": "Suzanne tessellated %dx%d as .obj text" % (o.n, o.n), "triangles": int(len(tris)), "vertices": int(len(pos)), "file_mb": size / 1e6,
           "host_threads": os.cpu_count(), "product": [], "reference": None, "mesh_cache": None}
    ref_model = None
    for t in [int(x) for x in o.threads.split(",")]:
        best, st_best, m = None, None, None
        for _ in range(3):
            st = {}
            t0 = time.perf_counter()
            m, _ = hostio.load_obj(path, "", threads=t, stats=st)
            dt = time.perf_counter() - t0
            if best is None or st["total_s"] < best:
                best, st_best = st["total_s"], st
        if ref_model is None:
            ref_model = m
        else:
            for k in ("pos", "nrm", "uv", "tris"):
                assert np.array_equal(m[k].view(np.uint32), ref_model[k].view(np.uint32)), ("thread count changed the result", t, k)
        out["product"].append({"threads": int(st_best["threads"]), "seconds": best, "mb_per_s": size / 1e6 / best, "mtris_per_s": len(tris) / 1e6 / best,
                               "phases_s": {k: st_best[k] for k in ("read_s", "scan_s", "resolve_s", "parse_s")}})
    assert np.array_equal(ref_model["tris"][:, :9], tris[:, :9])
    ref = None if o.no_reference else orc.ref()
    if ref is not None:
        t0 = time.perf_counter()
        h = ref.ref_load_obj(path.encode(), b"")
        dt = time.perf_counter() - t0
        sz = np.zeros(5, np.uint64)
        ref.ref_scene_sizes(h, orc.ptr(sz))
        rp, rn = np.zeros((int(sz[0]), 3), np.float32), np.zeros((int(sz[1]), 3), np.float32)
        ru, rt = np.zeros((int(sz[2]), 2), np.float32), np.zeros((int(sz[3]), 10), np.int32)
        ref.ref_scene_copy(h, orc.ptr(rp), orc.ptr(rn), orc.ptr(ru), orc.ptr(rt))
        ref.ref_scene_destroy(h)
        same = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in ((rp, ref_model["pos"]), (rn, ref_model["nrm"]), (ru, ref_model["uv"]), (rt, ref_model["tris"])))
        out["reference"] = {"kind": "fileloader.cpp + tinyobjloader 1.0.5 (oracle/_ref), 1 thread", "seconds": dt, "mb_per_s": size / 1e6 / dt, "same_arrays": bool(same)}
        assert same, "product loader != reference loader"
    cache = os.path.join(tmp, "tess%d.rastmesh" % o.n)
    hostio.save_mesh_cache(path, "", cache)
    t0 = time.perf_counter()
    h = hostio.lib().rasth_load_mesh_cache(cache.encode())
    dt = time.perf_counter() - t0
    hostio.lib().rasth_model_free(h)
    out["mesh_cache"] = {"file_mb": os.path.getsize(cache) / 1e6, "seconds": dt, "mb_per_s": os.path.getsize(cache) / 1e6 / dt}
    if o.cli:  # the whole command line on this scene: parse (or cache read), upload, one 4K frame, frame.png + depth.png
        import subprocess
        exe = os.path.join(ROOT, "rasteriser_b200", "renderer")
        lights = os.path.join(data, "threepoint.csv")
        runs = {}
        for name, extra in (("parse_obj", []), ("mesh_cache", ["--mesh-cache", cache])):
            t0 = time.perf_counter()
            r = subprocess.run([exe, "-o", path, "-l", lights, "-x", "3840", "-y", "2160", "--quiet", "--timing"] + extra, cwd=tmp, capture_output=True, text=True)
            runs[name] = {"seconds": time.perf_counter() - t0, "returncode": r.returncode,
                          "stages_s": {l.split("] ")[1].rsplit(" ", 2)[0]: float(l.split()[-2]) for l in r.stderr.splitlines() if l.startswith("[timing]")}}
            if r.returncode != 0:
                runs[name]["stderr"] = r.stderr[-300:]
        out["renderer_cli_4k_frame"] = runs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
