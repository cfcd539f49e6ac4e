"""The parts of the C ABI the parity tests do not reach: device-pointer outputs on a caller-owned stream,
profiling, launch counting, statistics.  Needs a B200: `-m gpu`."""
import numpy as np
import pytest

import orc
import scenes as S
from gpu_common import make_renderer, to_api_args
from rasteriser_b200 import api

pytestmark = pytest.mark.gpu


def test_device_outputs_on_a_torch_stream_equal_host_outputs():
    import torch
    r = make_renderer(S.scene("suzanne"), S.lights("threepoint"))
    try:
        W, H, n = 320, 240, 5
        poses = [api.Args(W, H, tait_bryan_angles=(0.1, api.spin_angle(0.3, k, n), 0.0)) for k in range(n)]
        host_frames, host_depths = r.draw_frames(poses, want_depth=True)
        stream = torch.cuda.Stream()
        r.set_stream(stream.cuda_stream)
        frames = torch.zeros((n, 3, H, W), dtype=torch.uint8, device="cuda")
        depths = torch.zeros((n, H, W), dtype=torch.float32, device="cuda")
        with torch.cuda.stream(stream):
            r.draw_frames_device(poses, frames.data_ptr(), depths.data_ptr())
            total = frames.sum()          # consumer ordered after the kernels on the same stream
        stream.synchronize()
        assert np.array_equal(frames.cpu().numpy(), host_frames)
        assert np.array_equal(depths.cpu().numpy().view(np.uint32), host_depths.view(np.uint32))
        assert int(total) == int(host_frames.astype(np.uint64).sum())
        # frames only (depths = NULL), then back to the context's own stream
        frames.zero_()
        r.draw_frames_device(poses, frames.data_ptr(), None)
        r.sync()
        assert np.array_equal(frames.cpu().numpy(), host_frames)
        r.use_own_stream()
        f1, _ = r.draw_frame(poses[2])
        assert np.array_equal(f1, host_frames[2])
        r.set_stream(0)                   # the legacy default stream is a valid choice too
        f2, _ = r.draw_frame(poses[3])
        assert np.array_equal(f2, host_frames[3])
    finally:
        r.close()


def test_device_draws_of_more_frames_than_one_batch():
    """Device-pointer draws run up to 240 frames per launch sequence, host-buffer draws 32: 250 small frames into device memory span two
    batches there (and eight on the host path) and must equal the host result frame by frame."""
    import torch
    r = make_renderer(S.scene("suzanne"), S.lights("threepoint"))
    try:
        W, H, n = 64, 48, 250
        poses = [api.Args(W, H, tait_bryan_angles=(0.05 * k, api.spin_angle(0.3, k, n), 0.0), scale=0.6 + 0.003 * k) for k in range(n)]
        host_frames, host_depths = r.draw_frames(poses, want_depth=True)
        frames = torch.full((n, 3, H, W), 7, dtype=torch.uint8, device="cuda")
        depths = torch.full((n, H, W), -1.0, dtype=torch.float32, device="cuda")
        for _ in range(2):
            r.draw_frames_device(poses, frames.data_ptr(), depths.data_ptr())
            r.sync()
            assert np.array_equal(frames.cpu().numpy(), host_frames)
            assert np.array_equal(depths.cpu().numpy().view(np.uint32), host_depths.view(np.uint32))
    finally:
        r.close()


def test_profiling_launch_count_and_stats():
    scene, lights = S.scene("suzanne"), S.lights("threepoint")
    r = make_renderer(scene, lights)
    try:
        oa = orc.make_args(640, 480)
        l0 = r.launch_count()
        r.set_profiling(True)
        r.draw_frame(to_api_args(oa))
        ms = r.pass_ms()
        r.set_profiling(False)
        assert set(ms) == {"clear", "vertex", "setup", "raster", "shade"}
        assert ms["vertex"] > 0 and ms["setup"] > 0 and ms["raster"] > 0 and ms["shade"] > 0 and sum(ms.values()) < 50
        assert r.launch_count() - l0 >= 4   # vertex, setup, raster, shade (+ clear when the slot is not known empty)
        st = r.stats()
        _, _, tri, cnt = orc.oracle_draw(scene, lights, oa, want_counters=True)
        assert st["triangles"] == 968 and st["front_facing"] == cnt.front_facing == 614
        assert st["visible_pixels"] == int((tri != orc.NO_TRIANGLE).sum()) == 109438
        assert st["queued_chunks"] > 0
    finally:
        r.close()


def test_depth_to_u8_matches_cimg_normalisation():
    r = make_renderer(S.scene("suzanne"), S.lights("threepoint"))
    try:
        for kw in (dict(), dict(disp=(0, 0, 2.0)), dict(disp=(5.0, 0, 0))):   # usual, negative NDC depth, empty frame (min == max)
            oa = orc.make_args(200, 120, **kw)
            _, depth = r.draw_frame(to_api_args(oa))
            want = np.zeros(depth.shape, np.uint8)
            orc.oracle().orc_depth_to_u8(orc.ptr(depth), depth.size, orc.ptr(want))
            assert np.array_equal(r.depth_to_u8(200, 120), want)
    finally:
        r.close()


def test_two_contexts_are_independent():
    a = make_renderer(S.scene("suzanne"), S.lights("threepoint"))
    b = make_renderer(S.scene("plane"), S.lights("normalmap"))
    try:
        fa, _ = a.draw_frame(api.Args(160, 120))
        fb, _ = b.draw_frame(api.Args(96, 64, tait_bryan_angles=(0.9, 0.3, 0.0)))
        fa2, _ = a.draw_frame(api.Args(160, 120))
        assert np.array_equal(fa, fa2) and fa.shape != fb.shape
        z = np.load(S.GOLDEN + "/small_frames.npz")
        assert np.array_equal(fa, z["suzanne_160x120_frame"])
    finally:
        a.close()
        b.close()


def test_shared_reciprocal_division_is_ieee_exact():
    """exact::div3 (three quotients by one divisor, the compiler's div.rn fast-path sequence with the reciprocal
    hoisted) against __fdiv_rn on 2^33 pseudo-random quotients incl. rounding-stress mantissas: no bit may differ."""
    r = api.Renderer(0)
    try:
        assert r.selftest_division(1 << 33, seed=12345) == 0
        assert r.selftest_division(1 << 30, seed=987654321) == 0
    finally:
        r.close()


def test_sparse_host_copy_delivers_the_same_bytes(monkeypatch):
    """Host-buffer draws copy only each frame's covered rectangle over PCIe and write the background on the host: the
    caller's buffers -- pre-filled with garbage here -- must end up byte-identical to whole-frame copies, for single
    frames, batches, bands, an empty frame and a frame that is covered edge to edge."""
    from rasteriser_b200 import api
    scene, lights = S.scene("suzanne"), S.lights("threepoint")
    poses = [api.Args(641, 483, tait_bryan_angles=(0.1 * k, 0.9 * k, 0.0), displacement=(0.3 * (k % 3) - 0.3, 0.0, 0.0)) for k in range(5)]
    poses.append(api.Args(641, 483, displacement=(40.0, 0.0, 0.0)))             # nothing on screen
    poses.append(api.Args(641, 483, scale=6.0, displacement=(0.0, 0.0, 1.0)))    # camera inside the model: covered edge to edge
    results = {}
    monkeypatch.setenv("RAST_SPARSE_MIN_BYTES", "0")  # (frames below 4 MB are copied whole by default)
    for sparse in ("0", "1"):
        monkeypatch.setenv("RAST_SPARSE_COPY", sparse)
        r = make_renderer(scene, lights)
        try:
            out = []
            b0 = r.d2h_bytes()
            for a in poses:  # single frames
                f, d = np.full((3, 483, 641), 0xAB, np.uint8), np.full((483, 641), -7.0, np.float32)
                r.draw_frame(a, f, d)
                out.append((f, d))
            fs, ds = np.full((len(poses), 3, 483, 641), 0xCD, np.uint8), np.full((len(poses), 483, 641), 5.0, np.float32)
            r.draw_frames(poses, fs, ds)  # one batch
            out.append((fs, ds))
            fs2 = np.full((len(poses), 3, 483, 641), 0xEF, np.uint8)
            r.draw_frames(poses, fs2, None)  # colour only
            out.append((fs2, None))
            r.set_band(100, 333)
            fb, db = np.full((3, 233, 641), 0x11, np.uint8), np.full((233, 641), 9.0, np.float32)
            r.draw_frame(poses[1], fb, db)
            out.append((fb, db))
            results[sparse] = (out, r.d2h_bytes() - b0)
        finally:
            r.close()
    whole, sparse = results["0"], results["1"]
    for (f0, d0), (f1, d1) in zip(whole[0], sparse[0]):
        assert np.array_equal(f0, f1)
        assert (d0 is None and d1 is None) or np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    f_empty, d_empty = sparse[0][5]
    assert not f_empty.any() and (d_empty == 1.0).all()
    assert sparse[1] < 0.9 * whole[1]  # fewer bytes crossed PCIe (two of the seven poses cover most of the frame)


def test_kept_visibility_slot_is_cleared_at_any_alignment():
    """rast_set_keep_visibility keeps the keys of a call's last frame; the next call clears that one slot.  With an odd pixel
    count an odd slot starts 8 bytes off a 16-byte boundary (k_clear stores 16 bytes at a time): 2, 3, 6 and 7 frames of
    641 x 483 and of 33 x 1, each drawn twice, must equal frame-by-frame draws."""
    from rasteriser_b200 import api
    r = make_renderer(S.scene("suzanne"), S.lights("threepoint"))  # keeps visibility
    one = make_renderer(S.scene("suzanne"), S.lights("threepoint"))
    try:
        for W, H in ((641, 483), (33, 1)):
            for n in (2, 3, 6, 7):
                poses = [api.Args(W, H, tait_bryan_angles=(0.2, 0.7 * k + 0.1 * n, 0.0)) for k in range(n)]
                want = [one.draw_frame(a) for a in poses]
                for _ in range(2):
                    fs, ds = r.draw_frames(poses, want_depth=True)
                    for k in range(n):
                        assert np.array_equal(fs[k], want[k][0]) and np.array_equal(ds[k].view(np.uint32), want[k][1].view(np.uint32)), (W, H, n, k)
    finally:
        r.close()
        one.close()


def test_page_locked_caller_buffers_and_the_drop_in_registry():
    """rast_host_register on buffers the caller owns: same bytes as a draw into pageable memory, registering twice is fine;
    the drop-in draw_frame page-locks the buffers it is handed, keeps at most PINNED_OUTPUTS_MAX of them and releases them
    in unpin_outputs()."""
    from rasteriser_b200 import api
    scene, lights = S.scene("suzanne"), S.lights("threepoint")
    r = make_renderer(scene, lights)
    try:
        a = api.Args(320, 240, tait_bryan_angles=(0.1, 0.6, 0.0))
        want_f, want_d = r.draw_frame(a)
        f, d = np.full((3, 240, 320), 0x77, np.uint8), np.full((240, 320), 3.0, np.float32)
        r.pin_host(f); r.pin_host(d); r.pin_host(f)
        for _ in range(3):
            r.draw_frame(a, f, d)
            assert np.array_equal(f, want_f) and np.array_equal(d.view(np.uint32), want_d.view(np.uint32))
        r.unpin_host(f); r.unpin_host(d)
        b0 = r.d2h_bytes()
        r.draw_frame(a, f, d)
        assert np.array_equal(f, want_f)
        assert r.d2h_bytes() - b0 == 320 * 240 * 7  # a frame this small comes back whole (no rectangle, no host-side fill)
    finally:
        r.close()
    import orc
    l10 = orc.lights_array(lights)
    try:
        for k in range(api.PINNED_OUTPUTS_MAX + 3):
            fb, db = np.empty((3, 240, 320), np.uint8), np.empty((240, 320), np.float32)
            api.draw_frame(scene.positions, scene.tris, scene.normals, scene.uvs, l10, scene.materials, a, fb, db)
            assert np.array_equal(fb, want_f) and np.array_equal(db.view(np.uint32), want_d.view(np.uint32))
            assert len(api._pinned_outputs) <= api.PINNED_OUTPUTS_MAX
    finally:
        api.invalidate()
        api.unpin_outputs()
    assert not api._pinned_outputs


class _Mapped:
    """numpy views of page-locked, device-mapped host memory (rast_host_alloc): the buffers k_deliver can store into."""
    def __init__(self):
        import ctypes as C
        from rasteriser_b200 import _lib
        self.C, self.lib, self.ptrs = C, _lib.load(), []

    def empty(self, shape, dtype, fill):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.lib.rast_host_alloc(n)
        assert p
        self.ptrs.append(p)
        a = np.ctypeslib.as_array(self.C.cast(p, self.C.POINTER(self.C.c_uint8)), (n,)).view(dtype).reshape(shape)
        a[...] = fill
        return a

    def free(self):
        for p in self.ptrs:
            self.lib.rast_host_free(self.C.c_void_p(p))
        self.ptrs = []


@pytest.mark.parametrize("size", [(640, 480), (1920, 1080), (48, 40), (160, 121)])
def test_zero_copy_delivery_writes_the_same_bytes_as_the_copy_engine(size, monkeypatch):
    """Page-locked mapped host buffers receive their covered row spans from k_deliver (no copy engine); the buffers -- pre-filled with
    garbage -- must end up byte-identical to the copy-engine path (RAST_DELIVER=0) and to pageable buffers, for single frames, a batch
    with an empty and a full-screen frame, colour only, a band, and with retained outputs on."""
    from rasteriser_b200 import api
    monkeypatch.setenv("RAST_SPARSE_MIN_BYTES", "0")
    W, H = size
    scene, lights = S.scene("suzanne"), S.lights("threepoint")
    poses = [api.Args(W, H, tait_bryan_angles=(0.1 * k, 0.9 * k, 0.0), displacement=(0.3 * (k % 3) - 0.3, 0.1 * k - 0.2, 0.0), scale=1.0 - 0.1 * k) for k in range(5)]
    poses.append(api.Args(W, H, displacement=(40.0, 0.0, 0.0)))            # nothing on screen
    poses.append(api.Args(W, H, scale=6.0, displacement=(0.0, 0.0, 1.0)))   # covered edge to edge
    monkeypatch.setenv("RAST_DELIVER", "0")
    ce = make_renderer(scene, lights)
    monkeypatch.setenv("RAST_DELIVER", "1")
    r = make_renderer(scene, lights)
    m = _Mapped()
    try:
        want_f, want_d = ce.draw_frames(poses, want_depth=True)
        fs, ds = m.empty((len(poses), 3, H, W), np.uint8, 0xAB), m.empty((len(poses), H, W), np.float32, -7.0)
        b0 = r.d2h_bytes()
        r.draw_frames(poses, fs, ds)
        moved = r.d2h_bytes() - b0
        assert np.array_equal(fs, want_f) and np.array_equal(ds.view(np.uint32), want_d.view(np.uint32))
        if W % 16 == 0:
            assert moved < 0.8 * (fs.nbytes + ds.nbytes)  # spans, not frames
        fs[...] = 0x5A
        r.draw_frames(poses, fs, None)  # colour only
        assert np.array_equal(fs, want_f)
        f1, d1 = m.empty((3, H, W), np.uint8, 0x33), m.empty((H, W), np.float32, 2.0)
        r.set_retained_outputs(True)
        for k in (0, 3, 5, 1, 6, 2):  # one pair of buffers redrawn, incl. empty and full-screen frames in between
            r.draw_frame(poses[k], f1, d1)
            assert np.array_equal(f1, want_f[k]) and np.array_equal(d1.view(np.uint32), want_d[k].view(np.uint32)), k
        r.set_retained_outputs(False)
        if H >= 40:
            y0, y1 = H // 5, H - H // 3
            r.set_band(y0, y1)
            fb, db = m.empty((3, y1 - y0, W), np.uint8, 0x11), m.empty((y1 - y0, W), np.float32, 9.0)
            r.draw_frame(poses[1], fb, db)
            assert np.array_equal(fb, want_f[1][:, y0:y1]) and np.array_equal(db.view(np.uint32), want_d[1][y0:y1].view(np.uint32))
    finally:
        r.close()
        ce.close()
        m.free()


def test_retained_outputs_rewrite_only_what_changes(monkeypatch):
    """rast_set_retained_outputs: the caller redraws into the buffers of the previous draw (the reference's spin loop does) and the
    library resets only the part of the previously covered rectangle that the new frame leaves.  Every draw must leave the
    buffers byte-identical to a draw without the promise -- moving, shrinking, vanishing and full-screen models, fewer frames
    than before, colour only, other buffers, a band."""
    from rasteriser_b200 import api
    monkeypatch.setenv("RAST_SPARSE_MIN_BYTES", "0")  # (frames below 4 MB are copied whole by default: nothing would be retained)
    scene, lights = S.scene("suzanne"), S.lights("threepoint")
    W, H = 641, 483
    def seq(k0):
        p = [api.Args(W, H, tait_bryan_angles=(0.1 * k, 0.9 * (k + k0), 0.0), displacement=(0.5 * ((k + k0) % 3) - 0.5, 0.2 * k0, 0.0), scale=1.0 - 0.15 * k0) for k in range(4)]
        p.append(api.Args(W, H, displacement=(40.0, 0.0, 0.0)) if k0 % 2 else api.Args(W, H, scale=6.0, displacement=(0.0, 0.0, 1.0)))  # nothing on screen / covered edge to edge
        p.append(api.Args(W, H, scale=0.3, displacement=(-0.8 + 0.4 * k0, 0.5, 0.0)))
        return p
    plain = make_renderer(scene, lights)
    r = make_renderer(scene, lights)
    try:
        r.set_retained_outputs(True)
        fs, ds = np.full((6, 3, H, W), 0xAB, np.uint8), np.full((6, H, W), -7.0, np.float32)
        moved = 0
        for k0 in range(4):
            b0 = r.d2h_bytes()
            r.draw_frames(seq(k0), fs, ds)
            moved = r.d2h_bytes() - b0
            want_f, want_d = plain.draw_frames(seq(k0), want_depth=True)
            assert np.array_equal(fs, want_f), "call %d" % k0
            assert np.array_equal(ds.view(np.uint32), want_d.view(np.uint32)), "call %d" % k0
        assert moved < 0.8 * fs.nbytes + ds.nbytes
        # fewer frames than the buffers hold: the rest keeps the previous call's frames
        r.draw_frames(seq(5)[:3], fs[:3], ds[:3])
        want_f3, want_d3 = plain.draw_frames(seq(5)[:3], want_depth=True)
        assert np.array_equal(fs[:3], want_f3) and np.array_equal(fs[3:], want_f[3:]) and np.array_equal(ds[:3].view(np.uint32), want_d3.view(np.uint32))
        r.draw_frames(seq(2), fs, ds)  # ... and all six again
        want_f, want_d = plain.draw_frames(seq(2), want_depth=True)
        assert np.array_equal(fs, want_f) and np.array_equal(ds.view(np.uint32), want_d.view(np.uint32))
        # colour only into the same colour buffer is another pair of buffers: a full draw (garbage must disappear), then retained again
        fs[:] = 0x5A
        r.draw_frames(seq(1), fs, None)
        assert np.array_equal(fs, plain.draw_frames(seq(1))[0])
        r.draw_frames(seq(3), fs, None)
        assert np.array_equal(fs, plain.draw_frames(seq(3))[0])
        # single frames through rast_draw_frame into one reused pair of buffers (the reference's loop), then a band
        f1, d1 = np.full((3, H, W), 0x33, np.uint8), np.full((H, W), 2.0, np.float32)
        for a in seq(0) + seq(1):
            r.draw_frame(a, f1, d1)
            wf, wd = plain.draw_frame(a)
            assert np.array_equal(f1, wf) and np.array_equal(d1.view(np.uint32), wd.view(np.uint32))
        r.set_band(100, 333)
        plain.set_band(100, 333)
        fb, db = np.full((3, 233, W), 0x11, np.uint8), np.full((233, W), 9.0, np.float32)
        for a in seq(2)[:3]:
            r.draw_frame(a, fb, db)
            wf, wd = plain.draw_frame(a)
            assert np.array_equal(fb, wf) and np.array_equal(db.view(np.uint32), wd.view(np.uint32))
        # switching the promise off: a full draw again
        r.set_retained_outputs(False)
        fb[:] = 0x77
        r.draw_frame(seq(2)[0], fb, db)
        assert np.array_equal(fb, plain.draw_frame(seq(2)[0])[0])
    finally:
        r.close()
        plain.close()


@pytest.mark.gpu
def test_drop_in_draw_frame_sees_edits_and_new_scenes():
    """The reference re-reads its vectors on every call (headers/drawing.h:16-18).  The drop-in keeps the scene on the GPU
    only while its content is unchanged: an edit in place, or a different scene in recycled arrays (same id()), must be drawn."""
    import orc
    import scenes as S
    from gpu_common import assert_parity
    sc = S.scene("suzanne")
    l7 = S.lights("threepoint")
    a = api.Args(160, 120)
    oa = orc.make_args(160, 120)

    def draw(pos):
        f, d = np.zeros((3, 120, 160), np.uint8), np.ones((120, 160), np.float32)
        api.draw_frame(pos, sc.tris, sc.normals, sc.uvs, orc.lights_array(l7), sc.materials, a, f, d)
        return f, d

    pos = np.array(sc.positions)
    f0, d0 = draw(pos)
    want0 = orc.oracle_draw(sc, l7, oa)
    assert np.array_equal(f0, want0[0]) and np.array_equal(d0.view(np.uint32), want0[1].view(np.uint32))
    pos *= np.float32(0.5)  # edited in place: same object, same address
    f1, d1 = draw(pos)
    sc2 = orc.Scene(pos, sc.normals, sc.uvs, sc.tris, sc.materials)
    want1 = orc.oracle_draw(sc2, l7, oa)
    assert np.array_equal(f1, want1[0]) and np.array_equal(d1.view(np.uint32), want1[1].view(np.uint32))
    assert not np.array_equal(f0, f1)
    api.invalidate()


def _draw_into_garbage(r, poses, W, H, depth=True):
    """Device-pointer draw into buffers pre-filled with garbage: every byte the caller sees must have been written by the
    shade pass (the tiles nothing was drawn into included)."""
    P = W * H
    n = len(poses)
    fdev, ddev = r.device_alloc(n * 3 * P + 64), r.device_alloc(n * 4 * P + 64)
    try:
        import ctypes as C
        lib = api._lib.load()
        # fill through a host staging copy (no torch in this test): draw once with junk, then overwrite with the pattern
        junk_f, junk_d = np.full(n * 3 * P, 0xAB, np.uint8), np.full(n * P, -7.0, np.float32)
        cudart = C.CDLL("libcudart.so")
        cudart.cudaMemcpy(C.c_void_p(fdev), junk_f.ctypes.data_as(C.c_void_p), C.c_size_t(junk_f.nbytes), 1)
        cudart.cudaMemcpy(C.c_void_p(ddev), junk_d.ctypes.data_as(C.c_void_p), C.c_size_t(junk_d.nbytes), 1)
        r.draw_frames_device(poses, fdev, ddev if depth else None)
        r.sync()
        f, d = np.empty((n, 3, H, W), np.uint8), np.empty((n, H, W), np.float32)
        r.device_read(fdev, f)
        r.device_read(ddev, d)
        return f, d
    finally:
        r.device_free(fdev)
        r.device_free(ddev)


@pytest.mark.parametrize("size", [(640, 480), (160, 120), (48, 40), (1920, 1080), (257, 129), (130, 33), (16, 16), (33, 1), (1, 70)])
def test_every_output_byte_is_written_untouched_tiles_included(size):
    """The shade pass skips the keys of 32 x 16 pixel tiles no pass drew into and writes the cleared frame there with wide
    stores (W % 16 == 0) or scalar ones: garbage pre-filled device outputs must equal the oracle everywhere, for image sizes on
    and off the tile grid, across calls (flags handed back), poses with nothing on screen and a different image shape of the
    same pixel count in between."""
    W, H = size
    scene, lights = S.scene("suzanne"), S.lights("threepoint")
    r = make_renderer(scene, lights)
    try:
        poses = [api.Args(W, H, tait_bryan_angles=(0.0, 0.7 * k, 0.1 * k), displacement=(0.8 * (k % 3) - 0.8, 0.0, 0.0)) for k in range(4)]
        poses.append(api.Args(W, H, displacement=(40.0, 0.0, 0.0)))  # nothing on screen
        for round_ in range(3):
            f, d = _draw_into_garbage(r, poses, W, H)
            for k, a in enumerate(poses):
                oa = orc.make_args(W, H, disp=a.displacement, angles=a.tait_bryan_angles)
                wf, wd, _ = orc.oracle_draw(scene, lights, oa)
                assert np.array_equal(f[k], wf), "frame %d round %d" % (k, round_)
                assert np.array_equal(d[k].view(np.uint32), wd.view(np.uint32)), "depth %d round %d" % (k, round_)
            if round_ == 0 and W != H:  # same pixel count, other shape: the flag layout changes under the same key buffer
                f2, d2 = _draw_into_garbage(r, [api.Args(H, W)], H, W)
                wf, wd, _ = orc.oracle_draw(scene, lights, orc.make_args(H, W))
                assert np.array_equal(f2[0], wf) and np.array_equal(d2[0].view(np.uint32), wd.view(np.uint32))
            poses = poses[::-1]
        ff, dd = _draw_into_garbage(r, poses[:2], W, H, depth=False)  # colour only (poses is reversed once more by now)
        assert np.array_equal(ff[0], f[len(poses) - 1]) and np.array_equal(ff[1], f[len(poses) - 2])
        assert (dd == -7.0).all()  # depths = NULL: the caller's depth memory is not written
    finally:
        r.close()


def test_output_frame_stride_interleaves_two_partitions():
    """rast_set_output_frame_stride: two calls (even frames, odd frames -- what ranks 0 and 1 of a 2-GPU run do) fill one
    sequence buffer that equals the dense single-call sequence; fnv1a64 of the product equals the test infrastructure's."""
    W, H, n = 320, 240, 9
    r = make_renderer(S.scene("suzanne"), S.lights("threepoint"))
    try:
        poses = [api.Args(W, H, tait_bryan_angles=(0.0, api.spin_angle(0.2, k, n), 0.0)) for k in range(n)]
        want_f, want_d = r.draw_frames(poses, want_depth=True)
        P = W * H
        fdev, ddev = r.device_alloc(n * 3 * P), r.device_alloc(n * 4 * P)
        r.set_output_frame_stride(2)
        for part in (0, 1):
            r.draw_frames_device(poses[part::2], fdev + part * 3 * P, ddev + part * 4 * P)
        r.set_output_frame_stride(1)
        r.sync()
        f, d = np.empty((n, 3, H, W), np.uint8), np.empty((n, H, W), np.float32)
        r.device_read(fdev, f)
        r.device_read(ddev, d)
        r.device_free(fdev)
        r.device_free(ddev)
        assert np.array_equal(f, want_f) and np.array_equal(d.view(np.uint32), want_d.view(np.uint32))
        assert api.fnv1a64(f) == orc.fnv(f) and api.fnv1a64(d[3]) == orc.fnv(d[3])
    finally:
        r.close()
