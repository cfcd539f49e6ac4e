"""Parity at BASELINE.json's workload shapes (SURVEY.md 8d generators, rasteriser_b200/synth.py): tessellated
Suzanne (tiny-triangle regime, configs 3/5), large overlapping triangles (high overdraw, config 4), 64 lights.
Sizes are chosen so the oracle finishes in seconds on the GPU box's host cores; the full-size config 3 mesh
(8 M triangles) is compared directly, the others through reduced instances plus size-independent properties
(band stitching, idempotence).  Needs a B200: `-m gpu`."""
import numpy as np
import pytest

import orc
import scenes as S
from gpu_common import assert_parity, gpu_draw, make_renderer, to_api_args
from rasteriser_b200 import synth

pytestmark = pytest.mark.gpu


def _tess_scene(n):
    base = S.scene("suzanne")
    pos, nrm, uv, tris = synth.tessellate(base.positions, base.normals, base.uvs, base.tris, n)
    return orc.Scene(pos, nrm, uv, tris, base.materials)


def test_config3_full_size_tessellated_suzanne_4k():
    """8 016 008 triangles at 3840x2160: every pixel's winning triangle, depth and colour vs the oracle."""
    scene = _tess_scene(91)
    assert len(scene.tris) == 8016008 and len(scene.positions) == 4141104
    lights = S.lights("threepoint")
    oa = orc.make_args(3840, 2160)
    r = make_renderer(scene, lights)
    try:
        got = gpu_draw(r, oa)
        want = orc.oracle_draw(scene, lights, oa, threads=1)  # one walk over 8 M triangles; bands would each walk them all
        assert_parity(got, want, "config 3")
        assert int((got[2] != orc.NO_TRIANGLE).sum()) == 2202146  # visible pixels recorded by the survey's probe of the reference
        st = r.stats()
        assert st["triangles"] == 8016008 and st["visible_pixels"] == 2202146
    finally:
        r.close()


def test_config5_shape_tessellated_64_lights():
    scene = _tess_scene(40)  # 1.55 M triangles
    lights = synth.random_lights(64)
    oa = orc.make_args(1920, 1080, angles=(0.2, 0.8, 0.0))
    r = make_renderer(scene, lights)
    try:
        assert_parity(gpu_draw(r, oa), orc.oracle_draw(scene, lights, oa, threads=1), "config 5 shape")
    finally:
        r.close()


@pytest.mark.parametrize("mode", ["auto", "chunk", "tile"])
def test_config4_shape_high_overdraw(mode, monkeypatch):
    """Large overlapping triangles, depth complexity ~50: exercises both raster schedules (chunk queue with early
    depth rejection; screen-tile bins, which "auto" switches to on the second call) and ties."""
    monkeypatch.setenv("RAST_RASTER_MODE", mode)
    W, H = 1920, 1080
    pos, nrm, uv, tris = synth.overdraw_scene(12000, W, H, radius_px=80.0)
    scene = orc.Scene(pos, nrm, uv, tris, [{"kd": (0.8, 0.8, 0.8), "texels": None}])
    lights = S.lights("threepoint")
    oa = orc.make_args(W, H)
    r = make_renderer(scene, lights)
    try:
        got = gpu_draw(r, oa)
        want, cnt = orc.oracle_draw(scene, lights, oa, threads=16, want_counters=True)[:3], None
        assert_parity(got, want, "config 4 shape")
        assert (got[2] != orc.NO_TRIANGLE).mean() > 0.99   # the frame is covered
        # size-independent properties at this shape: idempotence and band stitching
        again = gpu_draw(r, oa)
        assert all(np.array_equal(a, b) for a, b in zip(got, again))
        f, d, t = np.zeros_like(got[0]), np.zeros_like(got[1]), np.zeros_like(got[2])
        for g in range(8):
            y0, y1 = H * g // 8, H * (g + 1) // 8
            r.set_band(y0, y1)
            bf, bd = r.draw_frame(to_api_args(oa))
            f[:, y0:y1], d[y0:y1], t[y0:y1] = bf, bd, r.triangle_ids(W, y1 - y0)
        r.set_band(0, 0)
        assert np.array_equal(f, got[0]) and np.array_equal(d.view(np.uint32), got[1].view(np.uint32)) and np.array_equal(t, got[2])
    finally:
        r.close()


def test_queue_overflow_is_correct():
    """More work items than the queue holds: overflowing triangles are rasterised by their setup thread and
    the result is still exact; the queue then grows."""
    W, H = 7680, 4320
    pos, nrm, uv, tris = synth.overdraw_scene(3000, W, H, radius_px=2200.0, seed=7)  # ~5000 chunks per triangle => 15 M items
    scene = orc.Scene(pos, nrm, uv, tris, [{"kd": (0.5, 0.6, 0.7), "texels": None}])
    lights = S.lights("normalmap")
    oa = orc.make_args(W, H)
    r = make_renderer(scene, lights)
    try:
        got = gpu_draw(r, oa)          # first call overflows the initial queue
        st = r.stats()
        got2 = gpu_draw(r, oa)         # second call runs with the grown queue
        assert all(np.array_equal(a, b) for a, b in zip(got, got2))
        # oracle at this size would take minutes: compare a band of 64 rows instead (pixels are independent)
        want = orc.oracle_draw(scene, lights, oa, band=(2100, 2164))
        for a, b in zip(got, want):
            assert np.array_equal(a[..., 2100:2164, :], b[..., 2100:2164, :])
        assert st["queued_chunks"] > (1 << 23)
    finally:
        r.close()


def _large_case(name):
    import json
    import os
    return json.load(open(os.path.join(S.GOLDEN, "large_cases.json")))[name]


def test_config4_full_size_overdraw_8k_equals_reference_hashes():
    """200 000 triangles R = 80 px at 7680x4320 (depth complexity ~50, every pixel covered): frame and depth hashes the
    REFERENCE produced on this input (tests/golden/make_golden_large.py), winning-triangle ids of the oracle."""
    want = _large_case("config4_overdraw_8k")
    scene = orc.Scene(*synth.overdraw_scene(200000, 7680, 4320), [{"kd": (0.8, 0.8, 0.8), "texels": None}])
    assert len(scene.tris) == want["triangles"]
    r = make_renderer(scene, S.lights("threepoint"))
    try:
        f, d, t = gpu_draw(r, orc.make_args(7680, 4320))
        assert int((t != orc.NO_TRIANGLE).sum()) == want["visible_pixels"]
        assert orc.fnv(t) == want["tri_fnv"] and orc.fnv(d) == want["depth_fnv"] and orc.fnv(f) == want["frame_fnv"]
    finally:
        r.close()


def test_config5_full_size_50m_triangles_64_lights_equals_reference_hashes():
    """49 880 072 triangles, 64 lights, 3840x2160: same hashes as the reference's own frame on this input."""
    want = _large_case("config5_tess227_64lights_4k")
    scene = _tess_scene(227)
    assert len(scene.tris) == want["triangles"] == 49880072
    r = make_renderer(scene, synth.random_lights(64))
    try:
        f, d, t = gpu_draw(r, orc.make_args(3840, 2160))
        assert int((t != orc.NO_TRIANGLE).sum()) == want["visible_pixels"]
        assert orc.fnv(t) == want["tri_fnv"] and orc.fnv(d) == want["depth_fnv"] and orc.fnv(f) == want["frame_fnv"]
        assert r.stats()["triangles"] == 49880072
    finally:
        r.close()
