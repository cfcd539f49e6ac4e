"""Parity of the CUDA frame path (through the C ABI) with the CPU oracle.  Needs a B200: `-m gpu`."""
import numpy as np
import pytest

import orc
import scenes as S
from gpu_common import assert_parity, gpu_draw, make_renderer, to_api_args, ulp_distance
from rasteriser_b200 import api

pytestmark = pytest.mark.gpu

CASES = S.golden_cases()


@pytest.fixture(scope="module")
def renderers():
    cache = {}

    def get(scene_name, lights_name):
        key = (scene_name, lights_name)
        if key not in cache:
            cache[key] = make_renderer(S.scene(scene_name), S.lights(lights_name))
        return cache[key]
    yield get
    for r in cache.values():
        r.close()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_cases(case, renderers):
    """Every committed reference case: vs the oracle (ids, depth, colour) and vs the reference's own hashes."""
    oa = S.case_args(case)
    r = renderers(case["scene"], case["lights"])
    got = gpu_draw(r, oa)
    want = orc.oracle_draw(S.scene(case["scene"]), S.lights(case["lights"]), oa)
    assert_parity(got, want, case["name"])
    # the implementation is in fact bit-exact: its output hashes equal the reference's
    assert orc.fnv(got[1]) == case["depth_fnv"]
    assert orc.fnv(got[0]) == case["frame_fnv"]
    assert orc.fnv(r.depth_to_u8(case["width"], case["height"])) == case["depth_u8_fnv"]
    assert np.array_equal(r.light_trans_dirs().view(np.uint32).ravel(),
                          np.array([int(b, 16) for b in case["trans_dir_bits"]], np.uint32))


@pytest.mark.parametrize("seed,n_tris,size,cw", [(1, 8, (16, 16), False), (2, 40, (64, 48), False), (3, 120, (257, 129), True),
                                                 (4, 200, (97, 131), False), (5, 60, (1, 50), True), (6, 30, (33, 1), False),
                                                 (7, 150, (128, 128), True), (8, 90, (200, 150), False), (9, 1, (31, 17), False),
                                                 (10, 3000, (320, 200), False), (11, 20000, (512, 512), True)])
def test_random_soups(seed, n_tris, size, cw):
    scene = S.random_soup(seed, n_tris)
    lights = S.random_lights(seed, 1 + seed % 5)
    oa = orc.make_args(size[0], size[1], scale=0.9, disp=(0.05, -0.03, 0.2), angles=(0.1 * seed, 0.37 * seed, -0.2), wind_clockwise=cw)
    r = make_renderer(scene, lights)
    try:
        assert_parity(gpu_draw(r, oa), orc.oracle_draw(scene, lights, oa, threads=4), "soup %d" % seed)
    finally:
        r.close()


def test_tie_rule_duplicates():
    """Coplanar identical triangles: the lower index (first drawn) wins every covered pixel."""
    scene = S.random_soup(11, 8, with_uv=False)
    lights = S.random_lights(3, 2)
    oa = orc.make_args(160, 120)
    r = make_renderer(scene, lights)
    try:
        got = gpu_draw(r, oa)
        assert_parity(got, orc.oracle_draw(scene, lights, oa), "duplicates")
        assert not (got[2] == 1).any()
    finally:
        r.close()


def test_many_lights_and_missing_material():
    scene = S.random_soup(21, 300)
    scene.tris[::7, 9] = -1           # no material (untextured white)
    scene.tris[::5, 6:9] = -1         # faces without uvs
    lights = S.random_lights(5, 64)
    lights[:, 3] = 16.0
    oa = orc.make_args(300, 200, angles=(0.2, 0.4, 0.0))
    r = make_renderer(scene, lights)
    try:
        assert_parity(gpu_draw(r, oa), orc.oracle_draw(scene, lights, oa, threads=4), "64 lights")
    finally:
        r.close()


def test_small_triangles_dense_mesh():
    """A finely tessellated grid: most triangles cover 0-1 samples (the tiny-triangle path)."""
    n = 300
    xs = np.linspace(-1.1, 1.1, n + 1, dtype=np.float32)
    gx, gy = np.meshgrid(xs, xs)
    rng = np.random.RandomState(0)
    pos = np.stack([gx.ravel(), gy.ravel(), (0.3 * np.sin(3 * gx) * np.cos(2 * gy)).ravel().astype(np.float32)], 1).astype(np.float32)
    pos[:, :2] += rng.rand(len(pos), 2).astype(np.float32) * 0.002
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    v = np.concatenate([np.stack([a, b, c], 1), np.stack([c, b, d], 1)])
    tris = np.zeros((len(v), 10), np.int32)
    tris[:, 0:3] = v
    tris[:, 3:6] = v % 97
    tris[:, 6:9] = -1
    nrm = rng.randn(97, 3).astype(np.float32)
    scene = orc.Scene(pos, nrm, np.zeros((1, 2), np.float32), tris, [{"kd": (0.7, 0.6, 0.5), "texels": None}])
    lights = S.lights("threepoint")
    oa = orc.make_args(400, 300, angles=(0.3, 0.2, 0.1))
    r = make_renderer(scene, lights)
    try:
        assert_parity(gpu_draw(r, oa), orc.oracle_draw(scene, lights, oa, threads=4), "dense grid")
        st = r.stats()
        assert st["triangles"] == len(tris)
    finally:
        r.close()


def test_bands_equal_whole_frame(renderers):
    """Sort-first screen bands (rast_set_band) stitch to exactly the whole frame, for G = 2, 3, 8."""
    oa = orc.make_args(320, 243, angles=(0.2, 0.9, 0.1))
    r = renderers("suzanne", "threepoint")
    whole = gpu_draw(r, oa)
    want = orc.oracle_draw(S.scene("suzanne"), S.lights("threepoint"), oa)
    assert_parity(whole, want, "whole")
    H = oa.image_height
    try:
        for G in (2, 3, 8):
            f, d, t = np.zeros_like(whole[0]), np.zeros_like(whole[1]), np.zeros_like(whole[2])
            for g in range(G):
                y0, y1 = H * g // G, H * (g + 1) // G
                r.set_band(y0, y1)
                bf, bd = r.draw_frame(to_api_args(oa))
                bt = r.triangle_ids(oa.image_width, y1 - y0)
                assert bf.shape == (3, y1 - y0, oa.image_width)
                f[:, y0:y1], d[y0:y1], t[y0:y1] = bf, bd, bt
            assert np.array_equal(f, whole[0]) and np.array_equal(d.view(np.uint32), whole[1].view(np.uint32)) and np.array_equal(t, whole[2])
    finally:
        r.set_band(0, 0)


def test_batched_spin_frames_equal_single_frames(renderers):
    """rast_draw_frames (frames batched through the kernels together) == one rast_draw_frame per pose,
    and each equals the oracle; the spin schedule is a pure function of k."""
    r = renderers("suzanne", "threepoint")
    n, W, H = 37, 200, 150
    poses = [api.Args(W, H, tait_bryan_angles=(0.0, api.spin_angle(0.0, k, n), 0.0)) for k in range(n)]
    frames, depths = r.draw_frames(poses, want_depth=True)
    for k in (0, 1, 17, 36):
        f1, d1 = r.draw_frame(poses[k])
        assert np.array_equal(frames[k], f1) and np.array_equal(depths[k].view(np.uint32), d1.view(np.uint32))
        oa = orc.make_args(W, H, angles=(0.0, float(orc.oracle().orc_spin_angle(0.0, k, n)), 0.0))
        wf, wd, _ = orc.oracle_draw(S.scene("suzanne"), S.lights("threepoint"), oa)
        assert np.array_equal(frames[k], wf) and ulp_distance(depths[k], wd).max() <= 1
    frames2, _ = r.draw_frames(poses)
    assert np.array_equal(frames, frames2)


def test_draw_frame_function_mirrors_reference_signature():
    """api.draw_frame(vertices, faces, normals, uvs, lights, materials, args, frame, depth)."""
    sc = S.scene("plane")
    lights = orc.lights_array(S.lights("threepoint"))
    a = api.Args(96, 64, tait_bryan_angles=(0.9, 0.3, 0.0))
    frame, depth = np.zeros((3, 64, 96), np.uint8), np.ones((64, 96), np.float32)
    api.draw_frame(sc.positions, sc.tris, sc.normals, sc.uvs, lights, sc.materials, a, frame, depth)
    z = np.load(S.GOLDEN + "/small_frames.npz")
    assert np.array_equal(frame, z["plane_96x64_frame"])
    assert np.array_equal(depth.view(np.uint32), z["plane_96x64_depth"].view(np.uint32))
    assert np.abs(np.linalg.norm(lights[:, 7:10], axis=1) - 1).max() < 1e-6


def test_idempotent_and_deterministic(renderers):
    r = renderers("suzanne", "threepoint")
    a = api.Args(640, 480, tait_bryan_angles=(0.3, 1.0, 0.2), displacement=(0.0, 0.0, 2.2))
    f0, d0 = r.draw_frame(a)
    t0 = r.triangle_ids(640, 480)
    for _ in range(3):
        f, d = r.draw_frame(a)
        assert np.array_equal(f, f0) and np.array_equal(d.view(np.uint32), d0.view(np.uint32)) and np.array_equal(r.triangle_ids(640, 480), t0)


def test_nonfinite_vertices_do_not_crash(renderers):
    """--scale 4 puts a vertex on w = 0 (inf/NaN raster coordinates).  Unpinned corner: the
    device follows the same IEEE operations as the oracle, so results are compared but only reported."""
    r = renderers("suzanne", "threepoint")
    oa = orc.make_args(640, 480, scale=4.0)
    got = gpu_draw(r, oa)
    got2 = gpu_draw(r, oa)
    assert all(np.array_equal(a, b) for a, b in zip(got[::2], got2[::2]))  # deterministic
    want = orc.oracle_draw(S.scene("suzanne"), S.lights("threepoint"), oa, threads=4)
    mism = int((got[2] != want[2]).sum())
    print("scale 4 (w=0): %d / %d pixels with a different winner" % (mism, want[2].size))


def test_errors():
    r = api.Renderer(0)
    try:
        with pytest.raises(api.RastError):
            r.draw_frame(api.Args(64, 64))          # no mesh uploaded
        sc = S.scene("square")
        bad = sc.tris.copy()
        bad[0, 0] = 99
        with pytest.raises(api.RastError, match="out of range"):
            r.upload_mesh(sc.positions, bad, sc.normals, sc.uvs)
        r.upload_mesh(sc.positions, sc.tris, sc.normals, sc.uvs)
        r.upload_materials(sc.materials)
        r.set_lights(S.lights("threepoint"))
        with pytest.raises(api.RastError):
            r.draw_frame(api.Args(0, 64))
        f, d = r.draw_frame(api.Args(64, 64))
        assert f.any()
    finally:
        r.close()


def test_renderer_cli_writes_reference_images(tmp_path):
    """The C++ command line (reference flags) end to end: frame.png / depth.png decode to the reference's images."""
    import os
    import subprocess
    from PIL import Image
    from rasteriser_b200 import build
    exe = build.build_renderer()
    case = [c for c in CASES if c["name"] == "suzanne_640x480_pose1"][0]
    cmd = [exe, "-o", os.path.join(S.DATA, "Suzanne.obj"), "-l", os.path.join(S.DATA, "threepoint.csv"), "--mats-dir", S.DATA + "/",
           "-x", "640", "-y", "480", "--scale", "1.25", "--dx", "0.1", "--dy", "-0.2", "--dz", "0.5", "--rx", "0.3", "--ry", "1.0", "--rz", "0.2"]
    out = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "Loading 968 triangles..." in out.stdout and "Loaded model" in out.stdout and "Loaded texture" in out.stdout
    frame = np.ascontiguousarray(np.asarray(Image.open(tmp_path / "frame.png")).transpose(2, 0, 1))
    depth8 = np.ascontiguousarray(np.asarray(Image.open(tmp_path / "depth.png")))
    assert orc.fnv(frame) == case["frame_fnv"]
    assert orc.fnv(depth8) == case["depth_u8_fnv"]
    # headless spin: 5 frames, last one saved; equals a single frame at the same angle
    out = subprocess.run(cmd[:11] + ["-s", "--frames", "5", "--save-frames", "spin_%02u.png", "--record", "spin.png", "--record-delay", "50"],
                         cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "frames/s" in out.stdout and (tmp_path / "spin_04.png").exists()
    with Image.open(tmp_path / "spin.png") as rec:  # the recorded sequence: one animated PNG holding the same five frames
        assert rec.n_frames == 5 and rec.info.get("duration") == 50
        for k in (0, 4):
            rec.seek(k)
            single = np.asarray(Image.open(tmp_path / ("spin_%02d.png" % k)))
            assert np.array_equal(np.asarray(rec.convert("RGB")), single)
    oa = orc.make_args(640, 480, angles=(0.0, float(orc.oracle().orc_spin_angle(0.0, 4, 5)), 0.0))
    wf, _, _ = orc.oracle_draw(S.scene("suzanne"), S.lights("threepoint"), oa)
    got = np.ascontiguousarray(np.asarray(Image.open(tmp_path / "spin_04.png")).transpose(2, 0, 1))
    assert np.array_equal(got, wf)
    # --mesh-cache: the first run parses the .obj and writes the cache, the second reads it; same images
    cache = str(tmp_path / "suzanne.rastmesh")
    for run in range(2):
        out = subprocess.run(cmd + ["--mesh-cache", cache, "--load-threads", "3", "--frame-out", "c%d.png" % run, "--depth-out", "cd%d.png" % run],
                             cwd=tmp_path, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        assert os.path.exists(cache) and (("Loading 968 triangles..." in out.stdout) == (run == 0))
        assert orc.fnv(np.ascontiguousarray(np.asarray(Image.open(tmp_path / ("c%d.png" % run))).transpose(2, 0, 1))) == case["frame_fnv"]
        assert orc.fnv(np.ascontiguousarray(np.asarray(Image.open(tmp_path / ("cd%d.png" % run))))) == case["depth_u8_fnv"]
    # errors: missing -l, unreadable model
    assert subprocess.run([exe, "-o", "x.obj"], capture_output=True).returncode == 1
    assert subprocess.run([exe, "-o", "/nonexistent.obj", "-l", os.path.join(S.DATA, "threepoint.csv")], capture_output=True).returncode == 1


@pytest.fixture
def tile_mode(monkeypatch):
    """Force the screen-tile binned raster schedule (RAST_RASTER_MODE=tile; the default is the chunk queue)."""
    monkeypatch.setenv("RAST_RASTER_MODE", "tile")


@pytest.mark.parametrize("name", ["suzanne_640x480_pose1", "suzanne_640x480_dz2.2", "suzanne_257x129", "suzanne_1x37", "plane_640x480_threepoint", "suzanne_1920x1080"])
def test_tile_binned_schedule_golden(name, tile_mode):
    case = [c for c in CASES if c["name"] == name][0]
    r = make_renderer(S.scene(case["scene"]), S.lights(case["lights"]))
    try:
        got = gpu_draw(r, S.case_args(case))
        assert orc.fnv(got[0]) == case["frame_fnv"] and orc.fnv(got[1]) == case["depth_fnv"]
        assert_parity(got, orc.oracle_draw(S.scene(case["scene"]), S.lights(case["lights"]), S.case_args(case)), name + " (tile bins)")
    finally:
        r.close()


@pytest.mark.parametrize("seed,n_tris,size,cw", [(2, 40, (64, 48), False), (3, 120, (257, 129), True), (7, 150, (128, 128), True), (10, 3000, (320, 200), False)])
def test_tile_binned_schedule_soups_bands_and_batches(seed, n_tris, size, cw, tile_mode):
    scene = S.random_soup(seed, n_tris)
    lights = S.random_lights(seed, 3)
    oa = orc.make_args(size[0], size[1], scale=0.9, disp=(0.05, -0.03, 0.2), angles=(0.1 * seed, 0.37 * seed, -0.2), wind_clockwise=cw)
    want = orc.oracle_draw(scene, lights, oa, threads=4)
    r = make_renderer(scene, lights)
    try:
        assert_parity(gpu_draw(r, oa), want, "soup %d (tile bins)" % seed)
        # a band whose edges cut through tiles
        H = oa.image_height
        y0, y1 = H // 3, min(H, H // 3 + max(1, H // 2))
        r.set_band(y0, y1)
        bf, bd = r.draw_frame(to_api_args(oa))
        bt = r.triangle_ids(oa.image_width, y1 - y0)
        assert np.array_equal(bf, want[0][:, y0:y1]) and np.array_equal(bd.view(np.uint32), want[1][y0:y1].view(np.uint32)) and np.array_equal(bt, want[2][y0:y1])
        r.set_band(0, 0)
        # several frames in one batch
        frames, depths = r.draw_frames([to_api_args(oa)] * 3, want_depth=True)
        for k in range(3):
            assert np.array_equal(frames[k], want[0]) and np.array_equal(depths[k].view(np.uint32), want[1].view(np.uint32))
    finally:
        r.close()


@pytest.mark.parametrize("kw", [dict(), dict(wind_clockwise=True, angles=(0.2, 2.5, -0.4), disp=(0.3, 0.1, 0.9)), dict(scale=1.25, disp=(0.1, -0.2, 0.5), angles=(0.3, 1.0, 0.2))])
def test_flat_face_extension_matches_oracle(kw):
    """EXTENSION (not reference behaviour): rast_args.flat = RAST_FLAT_FACE shades with one normal per face.  The
    reference parses -f and ignores it, so there is nothing of the reference's to pin this against; the device must
    equal the oracle's definition, and flat = 1 must stay identical to flat = 0."""
    scene, lights = S.scene("suzanne"), S.lights("threepoint")
    r = make_renderer(scene, lights)
    try:
        smooth = orc.make_args(320, 240, **kw)
        ref_flag = orc.make_args(320, 240, flat=True, **kw)
        face = orc.make_args(320, 240, **kw)
        face.flat = 2
        a = to_api_args(smooth)
        f0, d0 = r.draw_frame(a)
        a1 = to_api_args(ref_flag)
        f1, d1 = r.draw_frame(a1)
        assert np.array_equal(f0, f1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))   # -f alone: no-op, like the reference
        a2 = api.Args(320, 240, scale=smooth.scale, displacement=tuple(smooth.displacement), tait_bryan_angles=tuple(smooth.tait_bryan_angles),
                      wind_clockwise=bool(smooth.wind_clockwise), flat=True, flat_mode="face")
        f2, d2 = r.draw_frame(a2)
        t2 = r.triangle_ids(320, 240)
        want = orc.oracle_draw(scene, lights, face)
        assert_parity((f2, d2, t2), want, "flat face")
        assert np.array_equal(d2.view(np.uint32), d0.view(np.uint32))     # visibility is unchanged
        assert (f2 != f0).any()                                            # the shading is not
    finally:
        r.close()


def test_flat_face_soup_and_cli(tmp_path):
    import os
    import subprocess
    from PIL import Image
    from rasteriser_b200 import build
    scene, lights = S.random_soup(4, 200), S.random_lights(4, 3)
    oa = orc.make_args(97, 131, angles=(0.4, 1.48, -0.2), wind_clockwise=True)
    oa.flat = 2
    r = make_renderer(scene, lights)
    try:
        a = to_api_args(oa)
        a.flat, a.flat_mode = True, "face"
        f, d = r.draw_frame(a)
        assert_parity((f, d, r.triangle_ids(97, 131)), orc.oracle_draw(scene, lights, oa), "flat soup")
    finally:
        r.close()
    exe = build.build_renderer()
    base = [exe, "-o", os.path.join(S.DATA, "Suzanne.obj"), "-l", os.path.join(S.DATA, "threepoint.csv"), "--mats-dir", S.DATA + "/", "-x", "200", "-y", "150", "--ry", "0.7"]
    for extra, name in ((["-f"], "ref.png"), (["-f", "--flat-mode", "face"], "face.png")):
        out = subprocess.run(base + extra + ["--frame-out", name, "--depth-out", "d_" + name], cwd=tmp_path, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
    ref_img = np.ascontiguousarray(np.asarray(Image.open(tmp_path / "ref.png")).transpose(2, 0, 1))
    face_img = np.ascontiguousarray(np.asarray(Image.open(tmp_path / "face.png")).transpose(2, 0, 1))
    sm = orc.make_args(200, 150, angles=(0.0, 0.7, 0.0))
    fa = orc.make_args(200, 150, angles=(0.0, 0.7, 0.0))
    fa.flat = 2
    assert np.array_equal(ref_img, orc.oracle_draw(S.scene("suzanne"), S.lights("threepoint"), sm)[0])   # -f == no -f, as in the reference
    assert np.array_equal(face_img, orc.oracle_draw(S.scene("suzanne"), S.lights("threepoint"), fa)[0])


def test_extension_texture_modulates_kd():
    """EXTENSION (not reference behaviour): has_texture = 1 | RAST_TEXTURE_MODULATE_KD makes the texel modulate the
    material's Kd (the reference drops Kd of a textured material, material.cpp:19-21).  Defined identically in oracle and device;
    with Kd = (1,1,1) it must equal the reference mode bit for bit."""
    base = S.scene("suzanne")
    lights = S.lights("threepoint")
    oa = orc.make_args(320, 240, angles=(0.1, 0.6, 0.0))
    frames = {}
    for kd in ((0.64, 0.3, 0.9), (1.0, 1.0, 1.0)):
        mats = [{"kd": kd, "texels": base.materials[0]["texels"], "modulate_kd": True}]
        scene = orc.Scene(base.positions, base.normals, base.uvs, base.tris, mats)
        r = make_renderer(scene, lights)
        try:
            got = gpu_draw(r, oa)
        finally:
            r.close()
        assert_parity(got, orc.oracle_draw(scene, lights, oa), "texture x Kd %s" % (kd,))
        assert orc.oracle_draw(scene, lights, oa)[0].tobytes() == got[0].tobytes()
        frames[kd] = got[0]
    ref_mode = orc.oracle_draw(base, lights, oa)[0]
    assert np.array_equal(frames[(1.0, 1.0, 1.0)], ref_mode) and not np.array_equal(frames[(0.64, 0.3, 0.9)], ref_mode)
