"""CPU check of rasteriser_b200/csrc/tight_bbox.h (the RAST_TIGHT_TINY kernel variant): every bbox pixel the rule drops
must be one the reference's literal barycentric test (drawing.cpp:41-49,111) rejects.  tests/tight_rule_check.c compiles
the SAME header the kernel includes and replays the literal test on random triangles in twelve regimes (sub-pixel,
slivers, nearly collinear, lattice-aligned, off-screen / huge, denormal, arbitrary bit patterns, ...).
RAST_TIGHT_TRIANGLES=<n> widens the sweep (1.8 G triangles were run once)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(os.environ.get("RAST_TIGHT_TRIANGLES", "1000000"))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("tight") / "tight_rule_check")
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "tight_rule_check.c"), "-lm"])
    return exe


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_dropped_pixels_are_rejected_by_the_literal_test(checker, seed):
    p = subprocess.run([checker, str(seed), str(N)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    r = json.loads(p.stdout)
    assert r["weak_violations"] == 0 and r["strong_violations"] == 0
    assert r["checked_pixels"] > N  # the sweep really exercised dropped pixels
    assert r["emptied"] > N // 10   # and whole triangles were dropped
    assert r["weakest_ratio"] > 2.0  # the proof's margin: a dropped pixel misses the candidate threshold by more than 2x


def test_rule_leaves_degenerate_input_alone(checker):
    """NaN / infinite / zero areas are the literal path's business: compile a tiny driver against the header."""
    src = r'''
#include <stdio.h>
#include <math.h>
#include "%s"
int main(void) {
    const float areas[] = {0.f, -0.f, INFINITY, NAN, 1e-45f};
    int bad = 0;
    for (int i = 0; i < 5; ++i) {
        uint32_t x0 = 10, y0 = 10, x1 = 11, y1 = 11;
        int r = rast_tight_bbox(10.3f, 10.3f, 10.8f, 10.4f, 10.5f, 10.8f, areas[i], &x0, &y0, &x1, &y1);
        bad += !(r == 1 && x0 == 10 && y0 == 10 && x1 == 11 && y1 == 11);
    }
    { /* a regular sub-pixel triangle between sample points vanishes */
        uint32_t x0 = 10, y0 = 10, x1 = 11, y1 = 11;
        const float area = fabsf((10.8f - 10.3f) * (10.8f - 10.3f) - (10.4f - 10.3f) * (10.5f - 10.3f));
        bad += rast_tight_bbox(10.3f, 10.3f, 10.8f, 10.4f, 10.5f, 10.8f, area, &x0, &y0, &x1, &y1) != 0;
    }
    { /* one that contains the sample point (11, 11) keeps exactly that pixel */
        uint32_t x0 = 10, y0 = 10, x1 = 12, y1 = 12;
        const float area = fabsf((11.6f - 10.6f) * (11.7f - 10.5f) - (10.7f - 10.5f) * (10.9f - 10.6f));
        int r = rast_tight_bbox(10.6f, 10.5f, 11.6f, 10.7f, 10.9f, 11.7f, area, &x0, &y0, &x1, &y1);
        bad += !(r == 1 && x0 == 11 && x1 == 11 && y0 == 11 && y1 == 11);
    }
    printf("%%d\n", bad);
    return bad;
}
''' % os.path.join(ROOT, "rasteriser_b200", "csrc", "tight_bbox.h")
    d = os.path.dirname(checker)
    c = os.path.join(d, "degenerate.c")
    with open(c, "w") as f:
        f.write(src)
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-o", os.path.join(d, "degenerate"), c, "-lm"])
    assert subprocess.run([os.path.join(d, "degenerate")], capture_output=True, text=True).stdout.strip() == "0"


def _raster_triangles(scene_pos, tris, width, height, angles=(0.0, 0.0, 0.0)):
    """Raster-space (x, y) of every triangle, with the reference's per-vertex operations (geometry.cpp:44-74) in numpy
    float32 (separately rounded, like the oracle) and the oracle's camera matrix."""
    import ctypes as C

    import numpy as np

    import orc
    lib = orc.oracle()
    a = orc.make_args(width, height, angles=angles)
    mv, cam, nm, view = (np.zeros(16, np.float32) for _ in range(4))
    lib.orc_frame_matrices.argtypes = [C.c_void_p] * 5
    lib.orc_frame_matrices(C.byref(a), orc.ptr(mv), orc.ptr(cam), orc.ptr(nm), orc.ptr(view))
    p = np.asarray(scene_pos, np.float32)
    x, y, z = p[:, 0:1], p[:, 1:2], p[:, 2:3]
    clip = (cam[0:4][None, :] * x + cam[4:8][None, :] * y) + (cam[8:12][None, :] * z + cam[12:16][None, :] * np.float32(1.0))
    ndx, ndy = clip[:, 0] / clip[:, 3], clip[:, 1] / clip[:, 3]
    rx = (np.float32(0.5) * (ndx + np.float32(1.0))) * np.float32(width)
    ry = (np.float32(0.5) * (-ndy + np.float32(1.0))) * np.float32(height)
    xy = np.stack([rx, ry], axis=1).astype(np.float32)
    return np.ascontiguousarray(xy[np.asarray(tris)[:, 0:3]])  # [T][3][2]


@pytest.mark.parametrize("n,width,height", [(24, 640, 480), (12, 3840, 2160)])
def test_rule_on_tessellated_suzanne(checker, tmp_path, n, width, height):
    """The mesh shape the variant is for (BASELINE configs 3 / 5): Suzanne subdivided n x n, here at sizes the CPU suite can
    afford; several poses.  Reports how much of the per-pixel work of the reference's bbox walk the rule removes."""
    import scenes as S
    from rasteriser_b200 import synth
    sc = S.scene("suzanne")
    pos, nrm, uv, tris = synth.tessellate(sc.positions, sc.normals, sc.uvs, sc.tris, n)
    for angles in [(0.0, 0.0, 0.0), (0.3, 1.1, -0.2), (0.0, 3.0, 0.0)]:
        path = str(tmp_path / "tris.bin")
        _raster_triangles(pos, tris, width, height, angles).tofile(path)
        p = subprocess.run([checker, "--file", path, str(width), str(height)], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        r = json.loads(p.stdout)
        assert r["triangles"] == len(tris) and r["weak_violations"] == 0 and r["strong_violations"] == 0
        if n == 24:  # sub-pixel triangles: more than half of the reference's pixel tests are provably idle
            assert r["dropped_pixels"] * 2 > r["bbox_pixels"] and r["emptied"] * 4 > r["triangles"]
