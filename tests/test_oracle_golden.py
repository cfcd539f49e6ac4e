"""The oracle against the golden fixtures produced by the reference itself (tests/golden/make_golden.py
ran the reference's unmodified hot path, oracle/_ref).  CPU only."""
import json
import os

import numpy as np
import pytest

import orc
import scenes as S

CASES = S.golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_frames(case):
    if case["width"] * case["height"] > 700 * 500 and "spin90" not in case["name"]:
        pass  # the 1080p cases take ~0.1 s each in the oracle: kept
    args = S.case_args(case)
    frame, depth, tri = orc.oracle_draw(S.scene(case["scene"]), S.lights(case["lights"]), args)
    assert orc.fnv(frame) == case["frame_fnv"]
    assert orc.fnv(depth) == case["depth_fnv"]
    assert int((depth != 1.0).sum()) == case["visible"]
    assert int((tri != orc.NO_TRIANGLE).sum()) == case["visible"]
    d8 = np.zeros(depth.shape, np.uint8)
    orc.oracle().orc_depth_to_u8(orc.ptr(depth), depth.size, orc.ptr(d8))
    assert orc.fnv(d8) == case["depth_u8_fnv"]


def test_oracle_matches_reference_full_arrays():
    z = np.load(os.path.join(S.GOLDEN, "small_frames.npz"))
    for case in CASES:
        if case["name"] + "_frame" not in z:
            continue
        frame, depth, _ = orc.oracle_draw(S.scene(case["scene"]), S.lights(case["lights"]), S.case_args(case))
        assert np.array_equal(frame, z[case["name"] + "_frame"])
        assert np.array_equal(depth.view(np.uint32), z[case["name"] + "_depth"].view(np.uint32))


def _bits(a):
    return ["%08x" % v for v in np.ascontiguousarray(a, np.float32).view(np.uint32).ravel()]


def test_oracle_known_answers():
    kat = json.load(open(os.path.join(S.GOLDEN, "kat.json")))
    lib = orc.oracle()
    spos = S.scene("suzanne").positions
    snrm = S.scene("suzanne").normals
    for p in kat["poses"]:
        a = orc.make_args(p["width"], p["height"], scale=p["scale"], disp=p["disp"], angles=p["angles"])
        mv, cam, nm, view = (np.zeros(16, np.float32) for _ in range(4))
        lib.orc_frame_matrices(orc.C.byref(a), orc.ptr(mv), orc.ptr(cam), orc.ptr(nm), orc.ptr(view))
        model = np.zeros(16, np.float32)
        lib.orc_transformation_matrix(a.scale, a.displacement, a.tait_bryan_angles, orc.ptr(model))
        assert _bits(model) == p["model"]
        assert _bits(view) == p["view"]
        assert _bits(cam) == p["camera"]
        assert _bits(nm) == p["normal_matrix"]
        rv = np.zeros((4, 4), np.float32)
        for i in range(4):
            lib.orc_raster_vertex(orc.ptr(cam), p["width"], p["height"], orc.ptr(spos[i]), orc.ptr(rv[i]))
        assert _bits(rv) == p["raster_v0_3"]
        cn = np.zeros(3, np.float32)
        lib.orc_transform_direction(orc.ptr(nm), orc.ptr(snrm[0]), orc.ptr(cn))
        assert _bits(cn) == p["camera_normal0"]
        assert _bits(np.float32(lib.orc_signed_area_2d(orc.ptr(rv[1]), orc.ptr(rv[0]), orc.ptr(rv[3])))) == p["signed_area_v1_v0_v3"]
    l10 = orc.lights_array(S.lights("threepoint"))
    view = np.zeros(16, np.float32)
    lib.orc_transformation_matrix(1.0, (orc.C.c_float * 3)(0, 0, -3), (orc.C.c_float * 3)(0, 0, 0), orc.ptr(view))
    lib.orc_transform_lights(orc.ptr(view), orc.ptr(l10), len(l10))
    assert _bits(l10[:, 7:10]) == kat["threepoint_trans_dir"]
    for s in kat["shade"]:
        n = np.array([int(b, 16) for b in s["normal"]], np.uint32).view(np.float32)
        alb = np.array([int(b, 16) for b in s["albedo"]], np.uint32).view(np.float32)
        out = np.zeros(3, np.uint32)
        lib.orc_shade(orc.ptr(n), orc.ptr(alb), orc.ptr(l10), len(l10), orc.ptr(out))
        assert [int(x) for x in out] == s["rgb"]
    sc = S.scene("suzanne").orc()
    for t in kat["texture"]:
        uv = np.array([int(b, 16) for b in t["uv"]], np.uint32).view(np.float32)
        out = np.zeros(3, np.float32)
        lib.orc_material_sample(orc.C.byref(sc.materials[0]), orc.ptr(uv), orc.ptr(out))
        assert _bits(out) == t["rgb"]


def test_survey_known_values():
    """Values recorded independently by the survey: counters and a few fp32 bit patterns."""
    a = orc.make_args(640, 480)
    frame, depth, tri, cnt = orc.oracle_draw(S.scene("suzanne"), S.lights("threepoint"), a, want_counters=True)
    assert (cnt.front_facing, cnt.bbox_tests, cnt.covered) == (614, 512318, 150840)
    assert int((tri != orc.NO_TRIANGLE).sum()) == 109438
    mv, cam, nm, view = (np.zeros(16, np.float32) for _ in range(4))
    orc.oracle().orc_frame_matrices(orc.C.byref(a), orc.ptr(mv), orc.ptr(cam), orc.ptr(nm), orc.ptr(view))
    assert _bits(cam)[0] == "3fe7c3b6" and _bits(cam)[5] == "401a8279" and _bits(cam)[10] == "bf8456c7" and _bits(cam)[14] == "40397dd3"
    rv = np.zeros(4, np.float32)
    orc.oracle().orc_raster_vertex(orc.ptr(cam), 640, 480, orc.ptr(S.scene("suzanne").positions[0]), orc.ptr(rv))
    assert _bits(rv) == ["434e8c81", "434574b9", "3f715ff7", "3ee52598"]


def test_threaded_and_banded_oracle_equal_single():
    sc, li = S.scene("suzanne"), S.lights("threepoint")
    a = orc.make_args(257, 129, angles=(0.1, 0.7, 0.0))
    f0, d0, t0 = orc.oracle_draw(sc, li, a)
    f1, d1, t1 = orc.oracle_draw(sc, li, a, threads=4)
    assert np.array_equal(f0, f1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32)) and np.array_equal(t0, t1)
    fb, db, tb = np.zeros_like(f0), np.ones_like(d0), np.full_like(t0, orc.NO_TRIANGLE)
    for y0, y1 in ((0, 40), (40, 41), (41, 129)):
        f, d, t = orc.oracle_draw(sc, li, a, band=(y0, y1))
        fb[:, y0:y1], db[y0:y1], tb[y0:y1] = f[:, y0:y1], d[y0:y1], t[y0:y1]
        assert (t[:y0] == orc.NO_TRIANGLE).all() and (t[y1:] == orc.NO_TRIANGLE).all()
    assert np.array_equal(f0, fb) and np.array_equal(d0.view(np.uint32), db.view(np.uint32)) and np.array_equal(t0, tb)


def test_flat_flag_is_a_noop_and_face_extension_differs():
    """-f is parsed and never read by the reference (arguments.cpp:45): flat = 1 must not change a single bit.
    flat = 2 is this repository's extension (face normals); it changes colours only, never visibility."""
    sc, li = S.scene("suzanne"), S.lights("threepoint")
    a0, a1, a2 = orc.make_args(160, 120), orc.make_args(160, 120, flat=True), orc.make_args(160, 120)
    a2.flat = 2
    f0, d0, t0 = orc.oracle_draw(sc, li, a0)
    f1, d1, t1 = orc.oracle_draw(sc, li, a1)
    f2, d2, t2 = orc.oracle_draw(sc, li, a2)
    assert np.array_equal(f0, f1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32)) and np.array_equal(t0, t1)
    assert orc.fnv(f0) == [c for c in CASES if c["name"] == "suzanne_160x120"][0]["frame_fnv"]
    assert np.array_equal(d0.view(np.uint32), d2.view(np.uint32)) and np.array_equal(t0, t2) and (f0 != f2).any()
    # on the flat square the face normal equals the (single) vertex normal: same image within rounding
    sq = S.scene("square")
    b0, b2 = orc.make_args(120, 90, angles=(0.3, 0.4, 0.1)), orc.make_args(120, 90, angles=(0.3, 0.4, 0.1))
    b2.flat = 2
    g0, _, _ = orc.oracle_draw(sq, li, b0)
    g2, _, _ = orc.oracle_draw(sq, li, b2)
    assert np.abs(g0.astype(int) - g2.astype(int)).max() <= 1
