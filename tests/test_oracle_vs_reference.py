"""The oracle against the reference's own compiled hot path (oracle/_ref/libref.so), live, on inputs
the committed fixtures do not cover: random triangle soups with shared edges, exact duplicates,
zero-area triangles, both windings, odd image sizes.  Skipped where the reference cannot exist
(GPU box: no /root/reference and no prebuilt libref.so).  CPU only."""
import os
import tempfile

import numpy as np
import pytest

import orc
import scenes as S

pytestmark = pytest.mark.skipif(orc.ref() is None, reason="oracle/_ref/libref.so not available")


def _ref_draw(scene, lights7, args, tmpdir):
    ref = orc.ref()
    paths, kd = [], []
    for i, m in enumerate(scene.materials):
        kd += list(m["kd"])
        t = m.get("texels")
        if t is None:
            paths.append(None)
        else:
            # the reference loads textures from files and min/max-normalises them (material.h:20-23);
            # write 8-bit texels whose normalisation reproduces the float texture exactly
            from PIL import Image
            q = np.round(t * 255).astype(np.uint8)
            p = os.path.join(tmpdir, "tex%d.ppm" % i)
            Image.fromarray(q.transpose(1, 2, 0)).save(p)
            paths.append(p.encode())
    arr = (orc.C.c_char_p * max(1, len(paths)))(*paths)
    kd = np.array(kd, np.float32)
    h = ref.ref_scene_create(orc.ptr(scene.positions), len(scene.positions), orc.ptr(scene.normals), len(scene.normals),
                             orc.ptr(scene.uvs), len(scene.uvs), orc.ptr(scene.tris), len(scene.tris), orc.ptr(kd), arr, len(scene.materials))
    assert h
    W, H = args.image_width, args.image_height
    f, d = np.zeros((3, H, W), np.uint8), np.zeros((H, W), np.float32)
    l10 = orc.lights_array(lights7)
    assert ref.ref_scene_draw(h, orc.ptr(l10), len(l10), W, H, args.scale, args.displacement, args.tait_bryan_angles, args.wind_clockwise, 0, orc.ptr(f), orc.ptr(d)) == 0
    ref.ref_scene_destroy(h)
    return f, d


def _quantised(scene):
    """Make the float textures exactly representable as normalised 8-bit so both sides sample identical texels."""
    for m in scene.materials:
        if m.get("texels") is not None:
            q = np.round(m["texels"] * 255).astype(np.float32)
            q[0, 0, 0], q[0, 0, 1] = 0.0, 255.0  # pin min/max so normalize(0,1) divides by exactly 255
            t = np.ascontiguousarray(q)
            orc.oracle().orc_normalize_texture(orc.ptr(t), t.size)
            m["texels"] = t
    return scene


@pytest.mark.parametrize("seed,n_tris,size,cw", [(1, 8, (16, 16), False), (2, 40, (64, 48), False), (3, 120, (257, 129), True),
                                                 (4, 200, (97, 131), False), (5, 60, (1, 50), True), (6, 30, (33, 1), False),
                                                 (7, 150, (128, 128), True), (8, 90, (200, 150), False)])
def test_random_soups(seed, n_tris, size, cw):
    scene = _quantised(S.random_soup(seed, n_tris))
    lights = S.random_lights(seed, 1 + seed % 4)
    args = orc.make_args(size[0], size[1], scale=0.9, disp=(0.05, -0.03, 0.2), angles=(0.1 * seed, 0.37 * seed, -0.2), wind_clockwise=cw)
    with tempfile.TemporaryDirectory() as tmp:
        rf, rd = _ref_draw(scene, lights, args, tmp)
    f, d, t = orc.oracle_draw(scene, lights, args)
    assert np.array_equal(f, rf)
    assert np.array_equal(d.view(np.uint32), rd.view(np.uint32))
    assert (t != orc.NO_TRIANGLE).sum() == (rd != 1.0).sum()


def test_duplicate_triangles_first_drawn_wins():
    """Two identical triangles with different materials: the strict '<' keeps the first (drawing.cpp:119)."""
    scene = S.random_soup(11, 8, degenerate=True, with_uv=False)
    lights = S.random_lights(3, 2)
    args = orc.make_args(80, 60)
    with tempfile.TemporaryDirectory() as tmp:
        rf, rd = _ref_draw(scene, lights, args, tmp)
    f, d, t = orc.oracle_draw(scene, lights, args)
    assert np.array_equal(f, rf) and np.array_equal(d.view(np.uint32), rd.view(np.uint32))
    assert not (t == 1).any()  # triangle 1 duplicates triangle 0 and must never win
