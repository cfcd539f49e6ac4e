"""Host side of the product (C++ loaders, PNG codec, flag parser) against the reference's own loader
output (tests/golden/scenes.npz came from fileloader.cpp + tinyobjloader) and, where the reference can be
compiled (dev container), against it live on adversarial OBJ files.  CPU only."""
import os
import shutil
import tempfile

import numpy as np
import pytest

from rasteriser_b200 import hostio as hostlib
import orc
import scenes as S

DATA = S.DATA


def test_suzanne_and_plane_equal_reference_loader():
    z = np.load(os.path.join(S.GOLDEN, "scenes.npz"))
    for name, obj in (("suzanne", "Suzanne.obj"), ("plane", "plane.obj")):
        m, warn = hostlib.load_obj(os.path.join(DATA, obj), DATA + "/")
        assert warn == ""
        for k, key in (("pos", "pos"), ("nrm", "nrm"), ("uv", "uv")):
            assert np.array_equal(m[k].view(np.uint32), z["%s_%s" % (name, key)].view(np.uint32)), (name, k)
        assert np.array_equal(m["tris"], z[name + "_tris"])
    m, _ = hostlib.load_obj(os.path.join(DATA, "Suzanne.obj"), DATA + "/")
    assert len(m["tris"]) == 968 and len(m["materials"]) == 1
    assert m["materials"][0]["kd"] == pytest.approx((0.64, 0.64, 0.64))
    # texture: same texels as the oracle-side reading (PIL decode + CImg-style normalisation)
    assert np.array_equal(m["materials"][0]["texels"].view(np.uint32), S.suzanne_texture().view(np.uint32))


def test_missing_mtl_is_a_warning_and_material_minus_one():
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copy(os.path.join(DATA, "plane.obj"), tmp)
        m, warn = hostlib.load_obj(os.path.join(tmp, "plane.obj"), os.path.join(tmp, "nowhere") + "/")
        assert "not found" in warn and len(m["materials"]) == 0
        assert (m["tris"][:, 9] == -1).all() and len(m["tris"]) == 2
    with pytest.raises(RuntimeError, match="Cannot open"):
        hostlib.load_obj("/nonexistent.obj")


def test_lights_csv():
    for name in ("threepoint", "normalmap"):
        assert np.array_equal(hostlib.load_lights(os.path.join(DATA, name + ".csv")), S.lights(name))
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "l.csv")
        open(p, "w").write("1,0,0,5,1,1,1\r\n\n0,1,0,6,0.5,0.5,0.5\n\n")   # CRLF, blank lines, trailing newline
        got = hostlib.load_lights(p)
        assert got.shape == (2, 7) and got[1, 3] == 6


ADVERSARIAL_OBJ = """# comment
mtllib adv.mtl
v 1 2 3
v -0.5 +0.25 1e2
v 1.5e-3 -2.5E+1 0.000000123456789
v 3.14159265358979 2.718281828 1.41421356
v 0.1 0.2 0.3 0.4
v 10 20
vn 0 0 1
vn 0.5773 0.5773 -0.5773
vt 0.25 0.75
vt 1.0 0.0 0.0
usemtl red
f 1/1/1 2/2/2 3/1/1
f 1//2 2//2 3//2 4//1 5//1
usemtl blue
f -1 -2 -3 -4
g second
f 1/2 2/1 3/2
usemtl nosuchmaterial
f 3 2 1
o third
usemtl red
f 1/1/1 3/2/2 5/1/2 6/2/1
"""
ADVERSARIAL_MTL = """newmtl red
Kd 0.9 0.1 0.05
Ka 1 1 1

newmtl blue
Kd 0.1 0.2 0.95
"""


@pytest.mark.skipif(orc.ref() is None, reason="oracle/_ref/libref.so not available")
def test_adversarial_obj_equals_reference_loader_live():
    """Number formats, relative indices, polygons, usemtl / g / o splitting: product loader == tinyobjloader as wrapped by the reference."""
    ref = orc.ref()
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "adv.obj"), "w").write(ADVERSARIAL_OBJ)
        open(os.path.join(tmp, "adv.mtl"), "w").write(ADVERSARIAL_MTL)
        m, warn = hostlib.load_obj(os.path.join(tmp, "adv.obj"), tmp + "/")
        h = ref.ref_load_obj(os.path.join(tmp, "adv.obj").encode(), (tmp + "/").encode())
        assert h
        sz = np.zeros(5, np.uint64)
        ref.ref_scene_sizes(h, orc.ptr(sz))
        pos, nrm = np.zeros((int(sz[0]), 3), np.float32), np.zeros((int(sz[1]), 3), np.float32)
        uv, tris = np.zeros((int(sz[2]), 2), np.float32), np.zeros((int(sz[3]), 10), np.int32)
        ref.ref_scene_copy(h, orc.ptr(pos), orc.ptr(nrm), orc.ptr(uv), orc.ptr(tris))
        ref.ref_scene_destroy(h)
    assert np.array_equal(m["pos"].view(np.uint32), pos.view(np.uint32))
    assert np.array_equal(m["nrm"].view(np.uint32), nrm.view(np.uint32))
    assert np.array_equal(m["uv"].view(np.uint32), uv.view(np.uint32))
    assert np.array_equal(m["tris"], tris)
    assert len(m["materials"]) == int(sz[4]) == 2


def test_float_parser_known_values():
    l = hostlib.lib()
    for s, want in (("1", 1.0), ("-0.437500", -0.4375), ("+3.5", 3.5), ("1e2", 100.0), ("2.5E-1", 0.25), ("abc", 0.0), ("-", 0.0), ("7.", 7.0)):
        assert l.rasth_parse_float(s.encode()) == np.float32(want), s


def test_png_roundtrip_against_pil():
    from PIL import Image
    rng = np.random.RandomState(0)
    l = hostlib.lib()
    with tempfile.TemporaryDirectory() as tmp:
        for c, (w, h) in ((3, (37, 21)), (1, (64, 5)), (3, (1, 1))):
            img = rng.randint(0, 256, (c, h, w)).astype(np.uint8)
            p = os.path.join(tmp, "t%d.png" % c)
            assert l.rasth_png_write(p.encode(), img.ctypes.data, w, h, c) == 0
            back = np.asarray(Image.open(p))
            assert np.array_equal(back if c == 1 else back.transpose(2, 0, 1), img[0] if c == 1 else img)
            # and the decoder reads a PIL-written file (filters, multiple IDAT chunks)
            q = os.path.join(tmp, "pil%d.png" % c)
            Image.fromarray(img[0] if c == 1 else img.transpose(1, 2, 0)).save(q, optimize=True)
            dims, out = np.zeros(3, np.uint32), np.zeros(w * h * c, np.uint8)
            assert l.rasth_png_read(q.encode(), dims.ctypes.data, out.ctypes.data, out.size) == 0
            assert tuple(dims) == (w, h, c)
            assert np.array_equal(out.reshape(h, w, c), img.transpose(1, 2, 0))


def test_flag_table_matches_reference():
    """arguments.cpp:15-33: names, defaults, required -l, aspect ratio, switches."""
    rc, a = hostlib.parse_args(["renderer", "-l", "threepoint.csv"])
    assert rc == 0 and (a["width"], a["height"]) == (540, 304) and a["obj"] == "" and a["scale"] == 1.0
    assert np.float32(a["aspect"]) == np.float32(540) / np.float32(304)
    rc, a = hostlib.parse_args("renderer -o Suzanne.obj -l t.csv --mats-dir sampledata/ -x 1920 -y 1080 -s -f --wind-clockwise --rx 0.1 --ry 0.2 --rz -0.3 --scale 2 --dx 1 --dy 2 --dz 3".split())
    assert rc == 0 and a["obj"] == "Suzanne.obj" and a["lights"] == "t.csv" and a["mats_dir"] == "sampledata/"
    assert (a["width"], a["height"], a["spin"], a["flat"], a["wind_clockwise"]) == (1920, 1080, True, True, True)
    assert a["angles"] == pytest.approx((0.1, 0.2, -0.3)) and a["disp"] == (1.0, 2.0, 3.0) and a["scale"] == 2.0
    rc, a = hostlib.parse_args("renderer --obj a.obj --lights b.csv --width 8 --height 4 --spin --flat".split())
    assert rc == 0 and (a["width"], a["height"], a["spin"], a["flat"]) == (8, 4, True, True)
    assert hostlib.parse_args(["renderer"])[0] == 3                       # -l is required
    assert hostlib.parse_args(["renderer", "-l", "a", "--bogus"])[0] == 3
    assert hostlib.parse_args(["renderer", "-l", "a", "-x", "abc"])[0] == 3
    assert hostlib.parse_args(["renderer", "-l", "a", "-x"])[0] == 3
    assert hostlib.parse_args(["renderer", "-l", "a", "-l", "b"])[0] == 3
    assert hostlib.parse_args(["renderer", "--help"])[0] == 1
    assert hostlib.parse_args(["renderer", "--version"])[0] == 2
    rc, a = hostlib.parse_args(["renderer", "-l", "a", "--", "--bogus"])
    assert rc == 0
