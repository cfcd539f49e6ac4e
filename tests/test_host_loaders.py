"""Host side of the product (C++ loaders, PNG codec, flag parser) against the reference's own loader
output (tests/golden/scenes.npz came from fileloader.cpp + tinyobjloader) and, where the reference can be
compiled (dev container), against it live on adversarial OBJ files.  CPU only."""
import ctypes as C
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


# ---- parallel OBJ reader (loaders at scale) -------------------------------------------------
def _ref_load(path, mats_dir):
    ref = orc.ref()
    h = ref.ref_load_obj(path.encode(), mats_dir.encode())
    assert h
    sz = np.zeros(5, np.uint64)
    ref.ref_scene_sizes(h, orc.ptr(sz))
    pos, nrm = np.zeros((int(sz[0]), 3), np.float32), np.zeros((int(sz[1]), 3), np.float32)
    uv, tris = np.zeros((int(sz[2]), 2), np.float32), np.zeros((int(sz[3]), 10), np.int32)
    ref.ref_scene_copy(h, orc.ptr(pos), orc.ptr(nrm), orc.ptr(uv), orc.ptr(tris))
    ref.ref_scene_destroy(h)
    return dict(pos=pos, nrm=nrm, uv=uv, tris=tris, n_materials=int(sz[4]))


def _same_model(a, b, what):
    for k in ("pos", "nrm", "uv"):
        assert a[k].shape == b[k].shape and np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (what, k)
    assert a["tris"].shape == b["tris"].shape and np.array_equal(a["tris"], b["tris"]), (what, "tris")


def _random_obj(rng, n_lines):
    """OBJ text exercising everything that carries state from line to line: number formats, relative indices,
    polygons and degenerate faces, usemtl (known / unknown / repeated), g / o (with and without pending faces),
    comments, blank lines, CRLF and bare-CR line ends."""
    out = ["mtllib adv.mtl"]
    nv = nvn = nvt = 0
    num = lambda: rng.choice(["%d" % rng.randint(-9, 10), "%.6f" % rng.uniform(-2, 2), "%+.3f" % rng.uniform(-2, 2), "%.9g" % rng.uniform(-1, 1),
                              "%.3e" % rng.uniform(-50, 50), "%.12f" % rng.uniform(0, 1), ".5", "7.", "1E2", "-0"])
    for _ in range(n_lines):
        r = rng.rand()
        if r < 0.30 or nv < 3:
            out.append("v %s %s %s" % (num(), num(), num())); nv += 1
        elif r < 0.38:
            out.append("vn %s %s %s" % (num(), num(), num())); nvn += 1
        elif r < 0.46:
            out.append("vt %s %s" % (num(), num())); nvt += 1
        elif r < 0.80:
            corners = []
            for _ in range(rng.choice([1, 2, 3, 3, 3, 3, 4, 4, 5, 7])):
                v = rng.randint(1, nv + 1) if rng.rand() < 0.7 else -rng.randint(1, nv + 1)
                style = rng.randint(0, 4)
                if style == 1 and nvt:
                    corners.append("%d/%d" % (v, rng.randint(1, nvt + 1)))
                elif style == 2 and nvt and nvn:
                    corners.append("%d/%d/%d" % (v, rng.randint(1, nvt + 1) if rng.rand() < 0.8 else -rng.randint(1, nvt + 1), rng.randint(1, nvn + 1)))
                elif style == 3 and nvn:
                    corners.append("%d//%d" % (v, -rng.randint(1, nvn + 1) if rng.rand() < 0.3 else rng.randint(1, nvn + 1)))
                else:
                    corners.append("%d" % v)
            out.append("f " + rng.choice([" ", "  ", "\t"]).join(corners) + rng.choice(["", " ", "  "]))
        elif r < 0.88:
            out.append("usemtl " + rng.choice(["red", "blue", "nosuchmaterial", "red"]))
        elif r < 0.94:
            out.append(rng.choice(["g ", "o "]) + rng.choice(["a", "b c", "thing"]))
        elif r < 0.97:
            out.append("# comment f 1 2 3")
        else:
            out.append("")
    text = ""
    for line in out:
        text += line + rng.choice(["\n", "\n", "\n", "\r\n"])
    return text


@pytest.mark.skipif(orc.ref() is None, reason="oracle/_ref/libref.so not available")
def test_random_objs_equal_reference_loader_for_any_piece_size_and_thread_count():
    l = hostlib.lib()
    rng = np.random.RandomState(1234)
    try:
        with tempfile.TemporaryDirectory() as tmp:
            open(os.path.join(tmp, "adv.mtl"), "w").write(ADVERSARIAL_MTL)
            for case in range(12):
                p = os.path.join(tmp, "r%d.obj" % case)
                text = _random_obj(rng, 40 + 60 * case)
                if case % 3 == 0:  # a size that is an exact multiple of the page size takes the read (not mmap) path
                    text += "#" * ((-len(text) - 1) % 4096) + "\n"
                    assert len(text) % 4096 == 0
                open(p, "w", newline="").write(text)
                want = _ref_load(p, tmp + "/")
                for piece, threads in ((0, 1), (1, 4), (37, 3), (256, 2), (4096, 8)):
                    l.rasth_set_obj_piece_bytes(piece)
                    got, _ = hostlib.load_obj(p, tmp + "/", threads=threads)
                    _same_model(got, want, "case %d piece %d threads %d" % (case, piece, threads))
                    assert len(got["materials"]) == want["n_materials"]
    finally:
        l.rasth_set_obj_piece_bytes(0)


def test_tessellated_mesh_roundtrip_threads_and_cache():
    """A 139 k-triangle OBJ written by synth.write_obj: the arrays do not depend on the thread count or the piece
    size, equal the reference loader's (when available), and survive the binary mesh cache unchanged."""
    from rasteriser_b200 import synth
    l = hostlib.lib()
    z = np.load(os.path.join(S.GOLDEN, "scenes.npz"))
    pos, nrm, uv, tris = synth.tessellate(z["suzanne_pos"], z["suzanne_nrm"], z["suzanne_uv"], z["suzanne_tris"], 12)
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copy(os.path.join(DATA, "Suzanne.mtl"), tmp)
        shutil.copy(os.path.join(DATA, "SuzanneTex.png"), tmp)
        p = os.path.join(tmp, "tess.obj")
        synth.write_obj(p, pos, nrm, uv, tris, mtllib="Suzanne.mtl")
        plain = os.path.join(tmp, "plain") + "/"   # the same material without its texture: the reference build cannot decode PNG here
        os.mkdir(plain)
        open(plain + "Suzanne.mtl", "w").write("newmtl Material\nKd 0.64 0.64 0.64\n")
        stats = {}
        base, _ = hostlib.load_obj(p, tmp + "/", threads=1, stats=stats)
        assert stats["file_bytes"] == os.path.getsize(p) and stats["threads"] == 1
        assert base["tris"].shape == tris.shape and np.array_equal(base["tris"][:, :9], tris[:, :9])
        assert np.allclose(base["pos"], pos, rtol=0, atol=1e-6)
        try:
            for piece, threads in ((0, 0), (100000, 5), (4097, 16)):
                l.rasth_set_obj_piece_bytes(piece)
                got, _ = hostlib.load_obj(p, tmp + "/", threads=threads)
                _same_model(got, base, "piece %d threads %d" % (piece, threads))
        finally:
            l.rasth_set_obj_piece_bytes(0)
        if orc.ref() is not None:
            _same_model(base, _ref_load(p, plain), "reference loader")
        c = os.path.join(tmp, "tess.rastmesh")
        hostlib.save_mesh_cache(p, tmp + "/", c)
        back = hostlib.load_mesh_cache(c)
        _same_model(back, base, "mesh cache")
        assert len(back["materials"]) == len(base["materials"]) == 1
        assert np.array_equal(back["materials"][0]["texels"], base["materials"][0]["texels"])
        with pytest.raises(RuntimeError):
            hostlib.load_mesh_cache(p)  # not a cache file


def test_png_writer_bands_decode_identically():
    """The writer deflates bands of rows on several threads into one zlib stream (one IDAT per band): PIL and the
    product's own decoder must read back the same pixels whatever the thread count."""
    from PIL import Image
    l = hostlib.lib()
    l.rasth_png_set_threads.argtypes = [C.c_uint]
    rng = np.random.RandomState(7)
    w, h = 1024, 771
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx // 4) % 256, (yy // 3) % 256, rng.randint(0, 256, (h, w))]).astype(np.uint8)  # gradients + noise
    try:
        with tempfile.TemporaryDirectory() as tmp:
            for threads in (1, 2, 3, 8, 0):
                l.rasth_png_set_threads(threads)
                for c, planes in ((3, img), (1, img[2:3])):
                    p = os.path.join(tmp, "b%d_%d.png" % (threads, c))
                    assert l.rasth_png_write(p.encode(), np.ascontiguousarray(planes).ctypes.data, w, h, c) == 0
                    with Image.open(p) as im:
                        im.load()  # a corrupt stream (bad Adler-32 / CRC) raises here
                        back = np.asarray(im)
                    assert np.array_equal(back if c == 1 else back.transpose(2, 0, 1), planes[0] if c == 1 else planes)
                    dims, out = np.zeros(3, np.uint32), np.zeros(w * h * c, np.uint8)
                    assert l.rasth_png_read(p.encode(), dims.ctypes.data, out.ctypes.data, out.size) == 0
                    assert np.array_equal(out.reshape(h, w, c), planes.transpose(1, 2, 0))
            sizes = [os.path.getsize(os.path.join(tmp, "b%d_3.png" % t)) for t in (1, 8)]
            assert sizes[1] < sizes[0] * 1.02  # cutting into bands costs almost nothing in size
    finally:
        l.rasth_png_set_threads(0)


def test_animated_png_frames_decode_identically():
    """--record: the spin sequence as one APNG (acTL / fcTL / fdAT around the writer's zlib streams); PIL must see
    every frame with the right pixels, and a plain PNG reader (the product's own) the first one."""
    from PIL import Image
    l = hostlib.lib()
    l.rasth_apng_write.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    rng = np.random.RandomState(11)
    n, w, h = 5, 320, 200
    frames = rng.randint(0, 256, (n, 3, h, w)).astype(np.uint8)
    frames[:, :, 50:150, 60:200] = 17  # compressible areas as well
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "spin.png")
        assert l.rasth_apng_write(p.encode(), frames.ctypes.data, n, w, h, 3, 40) == 0
        with Image.open(p) as im:
            assert getattr(im, "n_frames", 1) == n and im.info.get("duration") == 40
            for k in range(n):
                im.seek(k)
                assert np.array_equal(np.asarray(im.convert("RGB")).transpose(2, 0, 1), frames[k]), k
        dims, out = np.zeros(3, np.uint32), np.zeros(w * h * 3, np.uint8)
        assert l.rasth_png_read(p.encode(), dims.ctypes.data, out.ctypes.data, out.size) == 0
        assert np.array_equal(out.reshape(h, w, 3), frames[0].transpose(1, 2, 0))


def test_load_obj_shim_with_reference_signature(tmp_path):
    """include/rast_load_obj.hpp: the reference's load_obj signature (fileloader.h:16) on the parallel reader, compiled
    with stand-in types; same vectors as the reference's loader produced (golden scenes), same stdout lines."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "rasteriser_b200", "host")
    exe = str(tmp_path / "shim")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-o", exe, os.path.join(root, "tests", "shim_load_obj_main.cpp"),
                           os.path.join(host, "loaders.cpp"), os.path.join(host, "png.cpp"), "-lz", "-lpthread"])
    out = subprocess.run([exe, os.path.join(DATA, "Suzanne.obj"), DATA + "/"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith("Loaded texture ") and lines[1] == "Loading 968 triangles..." and lines[2].startswith("Loaded model ")
    f = lines[-1].split()
    z = np.load(os.path.join(S.GOLDEN, "scenes.npz"))

    def fnv(a):
        h = 1469598103934665603
        for b in np.ascontiguousarray(a).view(np.uint8).ravel().tolist():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return "%016x" % h
    assert [int(x) for x in f[1:6]] == [len(z["suzanne_pos"]), len(z["suzanne_nrm"]), len(z["suzanne_uv"]), len(z["suzanne_tris"]), 1]
    assert f[6] == fnv(z["suzanne_pos"]) and f[7] == fnv(z["suzanne_tris"]) and f[8].endswith("SuzanneTex.png")
    assert subprocess.run([exe, "/nonexistent.obj"], capture_output=True).returncode == 1   # fileloader.cpp:98-100
