"""Randomised parity sweep: random triangle soups and the sample scenes under random image sizes, poses (including
camera-inside-the-model and eye-plane crossings, where the reference draws without clipping), windings, light sets
and both raster schedules, each compared pixel by pixel (winning triangle, depth bits, colour) with the oracle.
RAST_FUZZ_SEEDS=<n> widens the sweep (default 24 seeds; 2 000 were run once per round on a B200)."""
import os

import numpy as np
import pytest

import orc
import scenes as S
from gpu_common import assert_parity, gpu_draw, make_renderer

pytestmark = pytest.mark.gpu

N_SEEDS = int(os.environ.get("RAST_FUZZ_SEEDS", "24"))
FIRST = int(os.environ.get("RAST_FUZZ_FIRST", "1000"))


def _case(seed):
    rng = np.random.RandomState(seed)
    kind = rng.choice(["soup", "soup", "soup_big", "suzanne", "plane"])
    if kind == "soup":
        scene = S.random_soup(seed, int(rng.choice([1, 2, 7, 30, 150, 600])), extent=float(rng.choice([0.3, 1.2, 4.0])), z_spread=float(rng.choice([0.05, 1.0, 3.5])),
                              n_materials=int(rng.choice([1, 2, 3])), with_uv=bool(rng.rand() < 0.7))
    elif kind == "soup_big":
        scene = S.random_soup(seed, int(rng.choice([3000, 12000])), extent=float(rng.choice([1.0, 2.5])), z_spread=1.5)
    else:
        scene = S.scene(kind)
    W, H = int(rng.choice([1, 2, 31, 64, 97, 160, 257, 320, 641])), int(rng.choice([1, 3, 33, 48, 120, 131, 200, 483]))
    near = rng.rand() < 0.25  # push the model through the eye plane / behind the camera
    disp = (float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-1.0, 1.0)), float(rng.uniform(2.0, 3.6) if near else rng.uniform(-2.5, 1.5)))
    oa = orc.make_args(W, H, scale=float(rng.choice([0.05, 0.5, 1.0, 1.7, 4.0])), disp=disp,
                       angles=tuple(float(a) for a in rng.uniform(-3.2, 3.2, 3)), wind_clockwise=bool(rng.rand() < 0.4))
    lights = S.random_lights(seed, int(rng.choice([1, 2, 3, 5, 9, 70])))
    if rng.rand() < 0.15:
        lights[0, 3] = -50.0  # a negative intensity: negative colour sums wrap through the unsigned cast like the reference's
    mode = str(rng.choice(["chunk", "chunk", "tile"]))
    return scene, lights, oa, mode, kind


@pytest.mark.parametrize("seed", range(FIRST, FIRST + N_SEEDS))
def test_fuzz(seed, monkeypatch):
    scene, lights, oa, mode, kind = _case(seed)
    monkeypatch.setenv("RAST_RASTER_MODE", mode)
    r = make_renderer(scene, lights)
    try:
        got = gpu_draw(r, oa)
        want = orc.oracle_draw(scene, lights, oa, threads=4)
        assert_parity(got, want, "fuzz seed %d (%s, %dx%d, %s)" % (seed, kind, oa.image_width, oa.image_height, mode))
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))  # exact, not only within the contract
    finally:
        r.close()
