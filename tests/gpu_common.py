"""Helpers for the GPU parity tests: run the product through its C ABI and compare with the oracle."""
import numpy as np

import orc
from rasteriser_b200 import api


def to_api_args(oa):
    return api.Args(oa.image_width, oa.image_height, scale=oa.scale, displacement=tuple(oa.displacement),
                    tait_bryan_angles=tuple(oa.tait_bryan_angles), wind_clockwise=bool(oa.wind_clockwise), flat=bool(oa.flat))


def make_renderer(scene, lights7, device=0):
    r = api.Renderer(device)
    r.upload_mesh(scene.positions, scene.tris, scene.normals, scene.uvs)
    r.upload_materials(scene.materials)
    r.set_lights(lights7)
    r.set_keep_visibility(True)  # the parity tests compare the winning triangle of every pixel
    return r


def ulp_distance(a, b):
    """Distance in units in the last place between two float32 arrays (same sign assumed or both tiny)."""
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def assert_parity(got, want, what=""):
    """got / want = (frame u8 [3,H,W], depth f32 [H,W], tri_id u32 [H,W]).
    Bar (BASELINE.json north_star): coverage and winning triangle ids bit-exact, depth within 1 ulp,
    colour within 1/255 per channel.  The implementation is expected to be exact on all three; the
    tolerances are the contract, the exact counts are reported on failure."""
    gf, gd, gt = got
    wf, wd, wt = want
    assert gt.shape == wt.shape, what
    bad = gt != wt
    assert not bad.any(), "%s: %d pixels with a different winning triangle (first at %s)" % (what, bad.sum(), np.argwhere(bad)[:1])
    ulp = ulp_distance(gd, wd)
    assert ulp.max() <= 1, "%s: depth differs by up to %d ulp at %d pixels" % (what, ulp.max(), (ulp > 1).sum())
    diff = np.abs(gf.astype(np.int16) - wf.astype(np.int16))
    assert diff.max() <= 1, "%s: colour differs by up to %d at %d samples" % (what, diff.max(), (diff > 1).sum())
    return int(ulp.max()), int(diff.max())


def gpu_draw(r, oa):
    a = to_api_args(oa)
    frame, depth = r.draw_frame(a)
    tri = r.triangle_ids(frame.shape[2], frame.shape[1])
    return frame, depth, tri
