"""Scene fixtures shared by the CPU and GPU tests (test infrastructure)."""
import json
import os

import numpy as np

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")
GOLDEN = os.path.join(HERE, "golden")

_cache = {}


def lights(name):
    return np.loadtxt(os.path.join(DATA, name + ".csv"), delimiter=",", dtype=np.float32).reshape(-1, 7)


def suzanne_texture():
    if "tex" not in _cache:
        _cache["tex"] = orc.load_texture_png(os.path.join(DATA, "SuzanneTex.png"))
    return _cache["tex"]


def scene(name):
    """'suzanne' | 'plane' | 'square' as loaded by the reference's own loader (tests/golden/scenes.npz)."""
    if name not in _cache:
        z = np.load(os.path.join(GOLDEN, "scenes.npz"))
        mats = {"suzanne": [{"kd": (0.64, 0.64, 0.64), "texels": suzanne_texture()}],
                "plane": [{"kd": (0.8, 0.8, 0.8), "texels": None}],
                "square": [{"kd": (1.0, 1.0, 1.0), "texels": None}]}[name]
        _cache[name] = orc.Scene(z[name + "_pos"], z[name + "_nrm"], z[name + "_uv"], z[name + "_tris"], mats)
    return _cache[name]


def golden_cases():
    return json.load(open(os.path.join(GOLDEN, "cases.json")))


def case_args(c):
    angles = np.array([int(b, 16) for b in c["angle_bits"]], np.uint32).view(np.float32)
    return orc.make_args(c["width"], c["height"], scale=c["scale"], disp=c["disp"], angles=[float(a) for a in angles], wind_clockwise=c["wind_clockwise"])


def random_soup(seed, n_tris, extent=1.2, z_spread=1.0, n_materials=2, with_uv=True, degenerate=True):
    """A random triangle soup around the origin (the camera looks down -z from z=+3).  Includes shared
    vertices (shared edges), exact duplicates (depth ties), zero-area triangles and integer-friendly
    coordinates."""
    rng = np.random.RandomState(seed)
    n_verts = max(3, n_tris)  # fewer vertices than 3*T => many shared edges
    pos = (rng.rand(n_verts, 3).astype(np.float32) * 2 - 1) * np.array([extent, extent, z_spread], np.float32)
    snap = rng.rand(n_verts) < 0.2
    pos[snap] = np.round(pos[snap] * 4) / 4  # snap some to a lattice
    nrm = rng.randn(max(1, n_verts // 2), 3).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = rng.rand(max(1, n_verts // 2), 2).astype(np.float32)
    tris = np.zeros((n_tris, 10), np.int32)
    tris[:, 0:3] = rng.randint(0, n_verts, (n_tris, 3))
    tris[:, 3:6] = rng.randint(0, len(nrm), (n_tris, 3))
    tris[:, 6:9] = rng.randint(0, len(uv), (n_tris, 3)) if with_uv else -1
    tris[:, 9] = rng.randint(0, n_materials, n_tris)
    if degenerate and n_tris >= 8:
        tris[1] = tris[0]                      # exact duplicate: first drawn must win
        tris[1, 9] = (tris[0, 9] + 1) % n_materials
        tris[3, 0:3] = tris[2, [1, 2, 0]]      # same triangle, rotated vertex order
        tris[4, 1] = tris[4, 0]                # zero area (two equal vertices)
        tris[5, 0:3] = tris[5, 0]              # a point
    mats = [{"kd": tuple(rng.rand(3).astype(np.float32)), "texels": None} for _ in range(n_materials)]
    if with_uv and n_materials > 1:
        t = rng.rand(3, 7, 5).astype(np.float32)  # tiny odd-sized texture, already in [0,1]
        mats[1]["texels"] = t
    return orc.Scene(pos, nrm, uv, tris, mats)


def random_lights(seed, n):
    rng = np.random.RandomState(seed)
    l = np.zeros((n, 7), np.float32)
    l[:, 0:3] = rng.randn(n, 3)
    l[:, 3] = rng.rand(n) * 300
    l[:, 4:7] = rng.rand(n, 3)
    return l
