"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol of
include/rast.h, its host math equals the oracle's bit for bit, and it refuses to run without a GPU
(no CPU fallback).  No compute entry point is exercised here."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import orc
import scenes as S
from rasteriser_b200 import _lib, api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "rast.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rast_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in include/rast.h"
    for name in sorted(declared):
        assert hasattr(lib, name), "librast_b200.so does not export " + name
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree: %s" % (declared ^ set(_lib.SYMBOLS))
    assert b"sm_100a" in lib.rast_version()


def test_struct_layouts_match_oracle_mirror():
    # the product ABI structs and the oracle's are declared independently; same field layout keeps tests honest
    for a, b in ((_lib.RastLight, orc.OrcLight), (_lib.RastMaterial, orc.OrcMaterial), (_lib.RastArgs, orc.OrcArgs)):
        assert C.sizeof(a) == C.sizeof(b)
        assert [(n, getattr(a, n).offset) for n, _ in a._fields_] == [(n, getattr(b, n).offset) for n, _ in b._fields_]


def _bits(a):
    return ["%08x" % v for v in np.ascontiguousarray(a, np.float32).view(np.uint32).ravel()]


def test_host_matrices_equal_oracle_and_reference(lib):
    kat = json.load(open(os.path.join(S.GOLDEN, "kat.json")))
    for p in kat["poses"]:
        a = api.Args(p["width"], p["height"], scale=p["scale"], displacement=p["disp"], tait_bryan_angles=p["angles"])
        mv, cam, nm, view = api.frame_matrices(a)
        assert _bits(cam) == p["camera"]
        assert _bits(nm) == p["normal_matrix"]
        assert _bits(view) == p["view"]
        assert _bits(mv) == p["modelview_oracle"]
    rng = np.random.RandomState(0)
    for _ in range(200):
        a = api.Args(int(rng.randint(1, 4000)), int(rng.randint(1, 3000)), scale=float(rng.rand() * 3 + 0.1),
                     displacement=tuple(rng.randn(3)), tait_bryan_angles=tuple(rng.randn(3) * 4))
        got = api.frame_matrices(a)
        oa = orc.make_args(a.image_width, a.image_height, a.scale, a.displacement, a.tait_bryan_angles)
        want = [np.zeros(16, np.float32) for _ in range(4)]
        orc.oracle().orc_frame_matrices(C.byref(oa), *[orc.ptr(w) for w in want])
        for g, w in zip(got, want):
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32))


def test_host_lights_and_spin_equal_oracle(lib):
    kat = json.load(open(os.path.join(S.GOLDEN, "kat.json")))
    l7 = S.lights("threepoint")
    arr = (_lib.RastLight * len(l7))()
    for i, row in enumerate(l7):
        arr[i].direction = (C.c_float * 3)(*row[:3])
    view = api.frame_matrices(api.Args(640, 480))[3]
    lib.rast_transform_lights(orc.ptr(view), arr, len(l7))
    assert _bits(np.array([list(x.trans_dir) for x in arr], np.float32)) == kat["threepoint_trans_dir"]
    for k in (0, 1, 90, 359, 719):
        assert np.float32(api.spin_angle(0.25, k, 720)).view(np.uint32) == np.float32(orc.oracle().orc_spin_angle(0.25, k, 720)).view(np.uint32)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product must fail loudly instead of computing anything."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present; the failure path is checked on CPU-only hosts")
    with pytest.raises(api.RastError, match="no CPU fallback"):
        api.Renderer(0)


def test_product_does_not_reference_the_oracle():
    """Nothing under rasteriser_b200/ or include/ may import, include or link oracle/."""
    for base in ("rasteriser_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dp, f), errors="ignore").read()
                    assert "liboracle" not in text and "oracle.h" not in text and "import orc" not in text, os.path.join(dp, f)


def test_draw_frame_rejects_wrong_output_buffers():
    """dtype / shape / contiguity of the caller's buffers are checked before anything reaches ctypes: a float64
    or strided buffer must raise, not be overrun.  The check precedes the context, so it runs without a GPU."""
    sc = S.scene("suzanne")
    l = S.lights("threepoint")
    a = api.Args(64, 48)
    good_f, good_d = np.zeros((3, 48, 64), np.uint8), np.ones((48, 64), np.float32)
    bad = [(np.zeros((3, 48, 64), np.float64), good_d), (np.zeros((48, 64, 3), np.uint8), good_d), (good_f, np.ones((48, 64), np.float64)),
           (np.zeros((3, 48, 128), np.uint8)[:, :, ::2], good_d), (good_f, np.ones((64, 48), np.float32).T), (good_f[:, :40], good_d)]
    for f, d in bad:
        with pytest.raises(api.RastError, match="must be a writeable C-contiguous"):
            api.draw_frame(sc.positions, sc.tris, sc.normals, sc.uvs, l, sc.materials, a, f, d)


def test_scene_fingerprint_sees_in_place_edits():
    sc = S.scene("suzanne")
    arrays = [np.array(sc.positions), np.array(sc.tris), np.array(sc.normals), np.array(sc.uvs)]
    k0 = api._fingerprint(arrays)
    assert k0 == api._fingerprint([a.copy() for a in arrays])  # same content at other addresses: same scene
    arrays[0][17, 1] = np.nextafter(arrays[0][17, 1], np.float32(9))
    assert api._fingerprint(arrays) != k0


def test_hash64_sees_every_byte_and_needs_no_gpu():
    """rast_hash64 (the drop-in's scene fingerprint) is a host function: equal buffers hash equal, any single-byte change, a
    changed length or a changed seed changes the hash -- for sizes around the 32-byte block and for an empty buffer."""
    import ctypes as C
    import numpy as np
    from rasteriser_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 31, 32, 33, 64, 1000, 4099):
        a = rng.integers(0, 256, n, dtype=np.uint8)
        h = lib.rast_hash64(a.ctypes.data if n else None, n, 1)
        c = a.copy()  # (kept alive across the call)
        assert h == lib.rast_hash64(c.ctypes.data if n else None, n, 1)
        assert h != lib.rast_hash64(a.ctypes.data if n else None, n, 2)
        for i in range(0, n, max(1, n // 13)):
            b = a.copy()
            b[i] ^= 0x40
            assert lib.rast_hash64(b.ctypes.data, n, 1) != h, (n, i)
        if n:
            assert lib.rast_hash64(a.ctypes.data, n - 1, 1) != h
    z = np.zeros(64, np.uint8)
    assert lib.rast_hash64(z.ctypes.data, 32, 1) != lib.rast_hash64(z.ctypes.data, 64, 1)  # zero padding is not invisible
