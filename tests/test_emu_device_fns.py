"""The kernels' per-thread functions, compiled for the HOST, against the oracle (CPU only, no GPU needed).

tests/emu_device_fns.cu drives the RAST_HD functions of rasteriser_b200/csrc/kernels.cuh -- raster_vertex, signed_area_2d,
bounding_box, tri_setup, edges / candidate / fragment, stage_item / raster_item (the warp rasteriser's inner loop, lane by
lane), shade_pixel, sample_texture and, by flag, the variants' rast_tight_bbox and prepare_triangle / shade_pixel_prep -- in
the order the kernels launch them.  The result must equal the oracle's bit for bit (winning triangle, depth bits, colour bytes).  This is how kernel-logic changes
are checked in the development container, which has no GPU; the -m gpu tests remain the parity tests proper."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import orc
import scenes as S
from rasteriser_b200 import _lib
from test_parity_gpu_fuzz import _case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
SRC = os.path.join(ROOT, "tests", "emu_device_fns.cu")
CSRC = os.path.join(ROOT, "rasteriser_b200", "csrc")
TIGHT, PRE_NORMALS, EARLY_Z, ALL_CHUNKS, FLAT_FACE, PREP, WARP, TILES = 1, 2, 4, 8, 16, 32, 64, 128


class EmuMaterial(C.Structure):
    _fields_ = [("kd", C.c_float * 3), ("has_texture", C.c_int32), ("tex_w", C.c_int32), ("tex_h", C.c_int32), ("texels", C.c_void_p)]


def load_emu(name="libemu.so", defines=()):
    """Build (when stale) and load build/<name>."""
    out = os.path.join(ROOT, "build", name)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-fmad=false",
                               "-Xcompiler", "-fPIC,-ffp-contract=off,-Wno-unknown-pragmas,-pthread", "-shared", "-o", out, SRC, "-lpthread"] + ["-D" + d for d in defines])
    lib = C.CDLL(out)
    lib.emu_draw.restype = C.c_int
    lib.emu_draw.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                             C.POINTER(EmuMaterial), C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def emu_draw(emu, scene, lights7, oa, flags, tiny_max=16, band=None):
    """One frame through the host-compiled device functions; matrices and light directions from the product's own host code
    (rast_frame_matrices / rast_transform_lights run without a GPU)."""
    rast = _lib.load()
    ra = _lib.RastArgs()
    ra.image_width, ra.image_height, ra.aspect_ratio, ra.scale = oa.image_width, oa.image_height, oa.aspect_ratio, oa.scale
    ra.displacement = oa.displacement
    ra.tait_bryan_angles = oa.tait_bryan_angles
    ra.wind_clockwise, ra.flat = oa.wind_clockwise, oa.flat
    mv, cam, nm, view = (np.zeros(16, np.float32) for _ in range(4))
    rast.rast_frame_matrices(C.byref(ra), orc.ptr(mv), orc.ptr(cam), orc.ptr(nm), orc.ptr(view))
    l10 = orc.lights_array(lights7)
    rast.rast_transform_lights(orc.ptr(view), l10.ctypes.data_as(C.POINTER(_lib.RastLight)), len(l10))
    mats = (EmuMaterial * max(1, len(scene.materials)))()
    for i, m in enumerate(scene.materials):
        mats[i].kd = (C.c_float * 3)(*m["kd"])
        t = m.get("texels")
        mats[i].has_texture = 0 if t is None else (3 if m.get("modulate_kd") else 1)
        if t is not None:
            mats[i].tex_h, mats[i].tex_w = t.shape[1], t.shape[2]
            mats[i].texels = t.ctypes.data
    W, H = oa.image_width, oa.image_height
    y0, y1 = band if band else (0, H)
    rows = y1 - y0
    rgb, depth, ids = np.zeros((3, rows, W), np.uint8), np.zeros((rows, W), np.float32), np.zeros((rows, W), np.uint32)
    rc = emu.emu_draw(orc.ptr(scene.positions), len(scene.positions), orc.ptr(scene.normals), len(scene.normals), orc.ptr(scene.uvs), len(scene.uvs),
                      orc.ptr(scene.tris), len(scene.tris), mats, len(scene.materials), orc.ptr(l10), len(l10), orc.ptr(cam), orc.ptr(nm), orc.ptr(mv),
                      int(oa.wind_clockwise), W, H, y0, y1, tiny_max, flags, orc.ptr(rgb), orc.ptr(depth), orc.ptr(ids))
    assert rc == 0
    return rgb, depth, ids


def assert_exact(got, want, what):
    assert np.array_equal(got[2], want[2]), "%s: %d pixels with a different winning triangle" % (what, int((got[2] != want[2]).sum()))
    assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), "%s: depth bits differ at %d pixels" % (what, int((got[1].view(np.uint32) != want[1].view(np.uint32)).sum()))
    assert np.array_equal(got[0], want[0]), "%s: colour differs at %d samples" % (what, int((got[0] != want[0]).sum()))


@pytest.mark.parametrize("flags", [PRE_NORMALS, 0, PRE_NORMALS | TIGHT, PRE_NORMALS | ALL_CHUNKS, PRE_NORMALS | ALL_CHUNKS | EARLY_Z, PREP, PREP | TIGHT])
def test_golden_cases_on_the_host(emu, flags):
    """Every golden case of tests/golden/cases.json small enough for the CPU suite: equal to the oracle AND to the hashes the
    reference itself produced."""
    n = 0
    for c in S.golden_cases():
        if c["width"] * c["height"] > 330000:
            continue
        scene, lights = S.scene(c["scene"]), S.lights(c["lights"])
        oa = S.case_args(c)
        got = emu_draw(emu, scene, lights, oa, flags)
        assert_exact(got, orc.oracle_draw(scene, lights, oa), "%s flags %d" % (c["name"], flags))
        assert orc.fnv(got[0]) == c["frame_fnv"] and orc.fnv(got[1]) == c["depth_fnv"], c["name"]
        n += 1
    assert n >= 10


@pytest.mark.parametrize("seed", range(1000, 1040))
def test_fuzz_cases_on_the_host(emu, seed):
    """The GPU fuzz distribution (tests/test_parity_gpu_fuzz.py::_case): soups through the eye plane, 1x1 images, both windings,
    up to 70 lights, negative intensities; tiny path with and without the tight-bbox rule, chunk path with and without early z."""
    scene, lights, oa, mode, kind = _case(seed)
    want = orc.oracle_draw(scene, lights, oa, threads=2)
    for flags, tiny in [(PRE_NORMALS, 16), (TIGHT, 64), (PRE_NORMALS | TIGHT, 1 << 30), (ALL_CHUNKS | EARLY_Z, 16), (PREP | TIGHT, 16)]:
        assert_exact(emu_draw(emu, scene, lights, oa, flags, tiny), want, "seed %d (%s) flags %d" % (seed, kind, flags))


def test_band_and_face_normals_on_the_host(emu):
    scene, lights = S.scene("suzanne"), S.lights("threepoint")
    oa = orc.make_args(320, 240, angles=(0.2, 0.7, 0.0))
    want = orc.oracle_draw(scene, lights, oa)
    got = emu_draw(emu, scene, lights, oa, PRE_NORMALS | TIGHT, band=(60, 180))
    assert_exact(got, tuple(a[..., 60:180, :] for a in want), "band")
    oa.flat = 2  # extension: face normals
    assert_exact(emu_draw(emu, scene, lights, oa, FLAT_FACE), orc.oracle_draw(scene, lights, oa), "flat face")


@pytest.fixture(scope="module")
def emu_blockz():
    """The same driver over the RAST_BLOCK_Z variant of raster_item (block-level depth rejection; a host "warp" is one lane, so
    the block's farthest stored depth is the lane's own four pixels' -- a different but equally valid rejection)."""
    return load_emu("libemu_blockz.so", ("RAST_BLOCK_Z=1",))


def test_block_level_depth_rejection_variant_on_the_host(emu_blockz):
    """High overdraw (the shape of BASELINE config 4: large overlapping triangles, depth complexity ~50, exact depth ties) plus
    golden and fuzz cases through the chunk rasteriser with early z: identical to the oracle."""
    from rasteriser_b200 import synth
    W, H = 640, 360
    pos, nrm, uv, tris = synth.overdraw_scene(4000, W, H, radius_px=60.0)
    scene = orc.Scene(pos, nrm, uv, tris, [{"kd": (0.8, 0.8, 0.8), "texels": None}])
    lights = S.lights("threepoint")
    oa = orc.make_args(W, H)
    want = orc.oracle_draw(scene, lights, oa, threads=4)
    assert (want[2] != orc.NO_TRIANGLE).mean() > 0.99
    assert_exact(emu_draw(emu_blockz, scene, lights, oa, PRE_NORMALS | ALL_CHUNKS | EARLY_Z), want, "overdraw, block z")
    assert_exact(emu_draw(emu_blockz, scene, lights, oa, PRE_NORMALS | TIGHT | EARLY_Z, 16), want, "overdraw, block z, tiny path mixed in")
    for c in S.golden_cases():
        if c["width"] * c["height"] > 330000:
            continue
        scene, lights = S.scene(c["scene"]), S.lights(c["lights"])
        oa = S.case_args(c)
        assert_exact(emu_draw(emu_blockz, scene, lights, oa, PRE_NORMALS | ALL_CHUNKS | EARLY_Z), orc.oracle_draw(scene, lights, oa), c["name"])
    for seed in range(1000, 1030):
        scene, lights, oa, mode, kind = _case(seed)
        assert_exact(emu_draw(emu_blockz, scene, lights, oa, ALL_CHUNKS | EARLY_Z), orc.oracle_draw(scene, lights, oa, threads=2), "seed %d" % seed)


def test_block_level_depth_rejection_is_exercised_and_a_wrong_bound_is_caught():
    """The same overdraw scene through a deliberately broken build (lower bound raised by 0.002 in NDC depth: blocks are rejected
    that should not be): the frame must differ -- i.e. the rejection path really runs in the test above, and the comparison sees it."""
    from rasteriser_b200 import synth
    broken = load_emu("libemu_blockz_broken.so", ("RAST_BLOCK_Z=1", "RAST_BLOCK_Z_TEST_BIAS=0.002f"))
    W, H = 640, 360
    pos, nrm, uv, tris = synth.overdraw_scene(4000, W, H, radius_px=60.0)
    scene = orc.Scene(pos, nrm, uv, tris, [{"kd": (0.8, 0.8, 0.8), "texels": None}])
    lights = S.lights("threepoint")
    oa = orc.make_args(W, H)
    want = orc.oracle_draw(scene, lights, oa, threads=4)
    got = emu_draw(broken, scene, lights, oa, PRE_NORMALS | ALL_CHUNKS | EARLY_Z)
    assert not np.array_equal(got[2], want[2])
    got = emu_draw(broken, scene, lights, oa, PRE_NORMALS | ALL_CHUNKS | TILES)  # the tile schedule's block test (block_behind) likewise
    assert not np.array_equal(got[2], want[2])


def test_block_level_depth_bound_brute_force(tmp_path):
    """tests/blockz_rule_check.cu: for random items staged with the kernel's own stage_item, every pixel the exact path accepts
    has z >= plane - M (the bound the block test relies on); the worst pixel uses a small fraction of the margin."""
    import json
    exe = str(tmp_path / "blockz_rule_check")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-fmad=false", "-Xcompiler", "-ffp-contract=off,-Wno-unknown-pragmas",
                           "-DRAST_BLOCK_Z=1", "-o", exe, os.path.join(ROOT, "tests", "blockz_rule_check.cu")])
    for seed in (1, 2):
        p = subprocess.run([exe, str(seed), "200000"], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        r = json.loads(p.stdout)
        assert r["violations"] == 0 and r["accepted_pixels"] > 1000000 and r["usable_items"] > 100000
        assert r["worst_margin_fraction"] < 0.25


def test_lockstep_warp_emulation(emu, emu_blockz):
    """WARP flag: the chunk rasteriser's 32 lanes run as 32 host threads in lockstep, with real votes (__any_sync) and, in the
    RAST_BLOCK_Z build, the real warp-wide maximum of the stored depths -- the decisions a device warp takes, not the
    one-lane-at-a-time stand-ins.  Overdraw scene (block rejection in every item), a golden scene and fuzz cases."""
    from rasteriser_b200 import synth
    W, H = 480, 270
    pos, nrm, uv, tris = synth.overdraw_scene(1200, W, H, radius_px=50.0)
    scene = orc.Scene(pos, nrm, uv, tris, [{"kd": (0.8, 0.8, 0.8), "texels": None}])
    lights = S.lights("threepoint")
    oa = orc.make_args(W, H)
    want = orc.oracle_draw(scene, lights, oa, threads=4)
    for lib in (emu, emu_blockz):
        assert_exact(emu_draw(lib, scene, lights, oa, PRE_NORMALS | ALL_CHUNKS | EARLY_Z | WARP), want, "overdraw, lockstep warp")
    c = [c for c in S.golden_cases() if c["name"] == "suzanne_160x120"][0]
    scene, lights, oa = S.scene(c["scene"]), S.lights(c["lights"]), S.case_args(c)
    for lib in (emu, emu_blockz):
        got = emu_draw(lib, scene, lights, oa, PRE_NORMALS | ALL_CHUNKS | EARLY_Z | WARP)
        assert orc.fnv(got[0]) == c["frame_fnv"] and orc.fnv(got[1]) == c["depth_fnv"]
    for seed in (1001, 1005, 1012):
        scene, lights, oa, mode, kind = _case(seed)
        if oa.image_width * oa.image_height > 100000 or len(scene.tris) > 2000:
            continue
        assert_exact(emu_draw(emu_blockz, scene, lights, oa, ALL_CHUNKS | EARLY_Z | WARP), orc.oracle_draw(scene, lights, oa, threads=2), "seed %d" % seed)


@pytest.mark.parametrize("flags", [PRE_NORMALS, PREP | TIGHT])
def test_extension_texture_modulates_kd_on_the_host(emu, flags):
    """The device's shade_pixel / shade_pixel_prep with has_texture = 1 | RAST_TEXTURE_MODULATE_KD (texel x Kd, an extension the
    reference does not have) against the oracle's definition; Kd = 1 must reproduce the reference mode."""
    base = S.scene("suzanne")
    lights = S.lights("threepoint")
    oa = orc.make_args(160, 120, angles=(0.1, 0.6, 0.0))
    scene = orc.Scene(base.positions, base.normals, base.uvs, base.tris, [{"kd": (0.64, 0.3, 0.9), "texels": base.materials[0]["texels"], "modulate_kd": True}])
    got = emu_draw(emu, scene, lights, oa, flags)
    assert_exact(got, orc.oracle_draw(scene, lights, oa), "texture x Kd")
    assert not np.array_equal(got[0], orc.oracle_draw(base, lights, oa)[0])
    white = orc.Scene(base.positions, base.normals, base.uvs, base.tris, [{"kd": (1.0, 1.0, 1.0), "texels": base.materials[0]["texels"], "modulate_kd": True}])
    assert np.array_equal(emu_draw(emu, white, lights, oa, flags)[0], orc.oracle_draw(base, lights, oa)[0])


def test_screen_tile_schedule_on_the_host(emu):
    """The high-overdraw path on the CPU: bins per 32 x 32 tile processed near to far, the tile's keys in a private array, every item's
    eight 16 x 8 blocks tested with the kernel's own block_behind against the farthest stored depths and the survivors rasterised by
    raster_item in its tile flavour (rectangle clipped to the tile, block grid anchored at the tile, early depth rejection against the
    tile's keys), then merged.  Frames must equal the oracle bit for bit: an overdraw scene (most blocks are rejected), golden scenes,
    fuzz seeds, a band whose edges cut through tiles, and tiny triangles mixed in."""
    from rasteriser_b200 import synth
    W, H = 640, 360
    pos, nrm, uv, tris = synth.overdraw_scene(4000, W, H, radius_px=60.0)
    over = orc.Scene(pos, nrm, uv, tris, [{"kd": (0.8, 0.8, 0.8), "texels": None}])
    lights = S.lights("threepoint")
    oa = orc.make_args(W, H)
    want = orc.oracle_draw(over, lights, oa, threads=4)
    assert_exact(emu_draw(emu, over, lights, oa, PRE_NORMALS | ALL_CHUNKS | TILES), want, "overdraw, tiles")
    assert_exact(emu_draw(emu, over, lights, oa, PRE_NORMALS | TIGHT | TILES, 16), want, "overdraw, tiles, tiny path mixed in")
    band = (37, 150)
    assert_exact(emu_draw(emu, over, lights, oa, PRE_NORMALS | ALL_CHUNKS | TILES, band=band), tuple(a[..., band[0]:band[1], :] for a in want), "overdraw band, tiles")
    for c in S.golden_cases():
        if c["width"] * c["height"] > 80000:
            continue
        scene, lts, ca = S.scene(c["scene"]), S.lights(c["lights"]), S.case_args(c)
        assert_exact(emu_draw(emu, scene, lts, ca, PRE_NORMALS | ALL_CHUNKS | TILES), orc.oracle_draw(scene, lts, ca), c["name"] + " (tiles)")
    for seed in range(1000, 1016):
        scene, lts, fa, mode, kind = _case(seed)
        if fa.image_width * fa.image_height > 100000 or len(scene.tris) > 4000:
            continue
        assert_exact(emu_draw(emu, scene, lts, fa, ALL_CHUNKS | TILES), orc.oracle_draw(scene, lts, fa, threads=2), "seed %d (tiles)" % seed)
