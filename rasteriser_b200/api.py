"""Host-side mirror of the reference's frame-path interface on top of the C ABI (include/rast.h).

The reference's boundary is one C++ function (headers/drawing.h:16-18):

    draw_frame(model_vertices, faces, model_vertnormals, vertuvs, lights, materials, arguments,
               frame_buffer, depth_buffer)

`draw_frame` below keeps that argument order and meaning with numpy arrays in place of the
std::vectors / CImg buffers; `Renderer` is the explicit form (upload once, draw many) that the spin
loop (renderer.cpp:94-126) and the benchmarks use.  Everything here forwards to librast_b200.so;
nothing is computed in Python."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import RastArgs, RastLight, RastMaterial, RastStats, RAST_PASS_NAMES, NO_TRIANGLE  # noqa: F401


class RastError(RuntimeError):
    pass


@dataclass
class Args:
    """struct Args (headers/arguments.h:7-21) with the reference's defaults (arguments.cpp:15-33)."""
    image_width: int = 540
    image_height: int = 304
    spin: bool = False
    flat: bool = False  # parsed by the reference, never read by the frame path
    flat_mode: str = "reference"  # extension: "face" with flat=True shades with one normal per face (RAST_FLAT_FACE)
    wind_clockwise: bool = False
    scale: float = 1.0
    displacement: tuple = (0.0, 0.0, 0.0)
    tait_bryan_angles: tuple = (0.0, 0.0, 0.0)  # rx, ry, rz
    obj_file: str = ""
    lights_file: str = ""
    materials_directory: str = ""

    @property
    def aspect_ratio(self):
        return float(np.float32(self.image_width) / np.float32(self.image_height))  # arguments.cpp:39

    def to_rast(self):
        a = RastArgs()
        a.image_width, a.image_height = int(self.image_width), int(self.image_height)
        a.aspect_ratio = self.aspect_ratio
        a.scale = float(self.scale)
        a.displacement = (C.c_float * 3)(*[float(x) for x in self.displacement])
        a.tait_bryan_angles = (C.c_float * 3)(*[float(x) for x in self.tait_bryan_angles])
        a.wind_clockwise = int(bool(self.wind_clockwise))
        a.flat = 2 if (self.flat and self.flat_mode == "face") else int(bool(self.flat))
        return a


@dataclass
class Material:
    """class Material (headers/material.h:11-25): diffuse colour, optional diffuse texture.
    texels: float32 [3, h, w], already normalised to [0,1] as the reference's constructor does."""
    kd: tuple = (1.0, 1.0, 1.0)
    texels: np.ndarray = field(default=None, repr=False)
    modulate_kd: bool = False  # extension (RAST_TEXTURE_MODULATE_KD): albedo = texel x Kd; the reference ignores Kd of a textured material


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _as_rast_args(a):
    return a if isinstance(a, RastArgs) else a.to_rast()


class Renderer:
    """One rast_ctx on one GPU."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.rast_create(int(device), C.byref(h))
        if rc != 0:
            raise RastError("rast_create failed (%d): %s" % (rc, (self._lib.rast_last_error(None) or b"").decode()))
        self._h = h
        self.device = int(device)
        self.n_lights = 0
        self._lights = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rast_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RastError("%s failed (%d): %s" % (what, rc, (self._lib.rast_last_error(self._h) or b"").decode()))

    # ---- scene ----
    def upload_mesh(self, model_vertices, faces, model_vertnormals, vertuvs):
        """faces: int32 [T,10] in struct Triangle order (headers/face.h:6-13)."""
        pos = np.ascontiguousarray(model_vertices, np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(model_vertnormals, np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(vertuvs, np.float32).reshape(-1, 2)
        tris = np.ascontiguousarray(faces, np.int32).reshape(-1, 10)
        self._check(self._lib.rast_upload_mesh(self._h, _ptr(pos), len(pos), _ptr(nrm), len(nrm), _ptr(uv), len(uv), _ptr(tris), len(tris)), "rast_upload_mesh")
        self.n_tris, self.n_vertices, self.n_normals = len(tris), len(pos), len(nrm)

    def upload_materials(self, materials):
        arr = (RastMaterial * max(1, len(materials)))()
        keep = []
        for i, m in enumerate(materials):
            kd = m.kd if isinstance(m, Material) else m["kd"]
            tex = m.texels if isinstance(m, Material) else m.get("texels")
            mod = m.modulate_kd if isinstance(m, Material) else m.get("modulate_kd", False)
            arr[i].kd = (C.c_float * 3)(*[float(x) for x in kd])
            if tex is not None:
                t = np.ascontiguousarray(tex, np.float32)
                keep.append(t)
                arr[i].has_texture, arr[i].tex_h, arr[i].tex_w, arr[i].texels = (3 if mod else 1), t.shape[1], t.shape[2], t.ctypes.data
        self._check(self._lib.rast_upload_materials(self._h, arr, len(materials)), "rast_upload_materials")

    def set_lights(self, lights):
        """lights: [L,7] rows of the lights CSV (dx,dy,dz,intensity,r,g,b) or [L,10] with trans_dir."""
        l = np.asarray(lights, np.float32)
        l = l.reshape(-1, l.shape[-1] if l.ndim > 1 else 7)
        arr = (RastLight * max(1, len(l)))()
        for i, row in enumerate(l):
            arr[i].direction = (C.c_float * 3)(*row[0:3])
            arr[i].intensity = float(row[3])
            arr[i].colour = (C.c_float * 3)(*row[4:7])
        self._check(self._lib.rast_set_lights(self._h, arr, len(l)), "rast_set_lights")
        self.n_lights, self._lights = len(l), arr

    def set_band(self, y0, y1):
        self._check(self._lib.rast_set_band(self._h, int(y0), int(y1)), "rast_set_band")
        self._band = (int(y0), int(y1)) if y1 > y0 else None

    def set_output_plane_stride(self, pixels):
        """Device-pointer draws write planes `pixels` apart (0 = one band): a band rendered straight into its rows of a full-size image."""
        self._check(self._lib.rast_set_output_plane_stride(self._h, int(pixels)), "rast_set_output_plane_stride")

    def set_output_frame_stride(self, frames):
        """Device-pointer draws put the i-th frame of a call into slot i * frames (1 = dense): N ranks fill one sequence buffer round-robin."""
        self._check(self._lib.rast_set_output_frame_stride(self._h, int(frames)), "rast_set_output_frame_stride")

    # ---- peer memory (one process per GPU; include/rast.h "peer memory") ----
    def device_alloc(self, nbytes):
        p = self._lib.rast_device_alloc(self._h, int(nbytes))
        if not p:
            raise RastError("rast_device_alloc: " + (self._lib.rast_last_error(self._h) or b"").decode())
        return int(p)

    def device_free(self, ptr):
        self._check(self._lib.rast_device_free(self._h, C.c_void_p(int(ptr))), "rast_device_free")

    def device_read(self, ptr, out):
        """Copy out.nbytes bytes from device address `ptr` into the numpy array `out`."""
        self._check(self._lib.rast_device_read(self._h, out.ctypes.data, C.c_void_p(int(ptr)), out.nbytes), "rast_device_read")
        return out

    def ipc_export(self, ptr):
        h = (C.c_ubyte * 64)()
        self._check(self._lib.rast_ipc_export(self._h, C.c_void_p(int(ptr)), h), "rast_ipc_export")
        return bytes(h)

    def ipc_open(self, handle):
        h = (C.c_ubyte * 64)(*handle)
        p = C.c_void_p()
        self._check(self._lib.rast_ipc_open(self._h, h, C.byref(p)), "rast_ipc_open")
        return int(p.value)

    def ipc_close(self, ptr):
        self._check(self._lib.rast_ipc_close(self._h, C.c_void_p(int(ptr))), "rast_ipc_close")

    def set_stream(self, cuda_stream):
        """Launch on this cudaStream_t handle (int; 0 = the legacy default stream)."""
        self._check(self._lib.rast_set_stream(self._h, C.c_void_p(int(cuda_stream))), "rast_set_stream")

    def use_own_stream(self):
        self._check(self._lib.rast_use_own_stream(self._h), "rast_use_own_stream")

    # ---- drawing ----
    def draw_frame(self, args, frame=None, depth=None, want_depth=True):
        """draw_frame into host arrays (allocated if not given). Returns (frame u8 [3,h,W], depth f32 [h,W] | None)."""
        a = _as_rast_args(args)
        rows = self._band_rows(a.image_height)
        if frame is None:
            frame = np.empty((3, rows, a.image_width), np.uint8)
        if depth is None and want_depth:
            depth = np.empty((rows, a.image_width), np.float32)
        _check_out("frame", frame, np.uint8, (3, rows, a.image_width))
        if depth is not None:
            _check_out("depth", depth, np.float32, (rows, a.image_width))
        self._check(self._lib.rast_draw_frame(self._h, C.byref(a), _ptr(frame), _ptr(depth), self._lights), "rast_draw_frame")
        return frame, depth

    def draw_frames(self, args_list, frames=None, depths=None, want_depth=False):
        """Several frames of the uploaded scene in one call (host outputs)."""
        n = len(args_list)
        arr = args_list if isinstance(args_list, C.Array) else (RastArgs * n)(*[_as_rast_args(a) for a in args_list])
        rows = self._band_rows(arr[0].image_height)
        if frames is None:
            frames = np.empty((n, 3, rows, arr[0].image_width), np.uint8)
        if depths is None and want_depth:
            depths = np.empty((n, rows, arr[0].image_width), np.float32)
        _check_out("frames", frames, np.uint8, (n, 3, rows, arr[0].image_width))
        if depths is not None:
            _check_out("depths", depths, np.float32, (n, rows, arr[0].image_width))
        self._check(self._lib.rast_draw_frames(self._h, arr, n, _ptr(frames), _ptr(depths), 0), "rast_draw_frames")
        return frames, depths

    def draw_frames_device(self, args_list, frames_ptr, depths_ptr=None):
        """Several frames into DEVICE memory (raw pointers, e.g. torch tensor .data_ptr()); asynchronous."""
        n = len(args_list)
        arr = args_list if isinstance(args_list, C.Array) else (RastArgs * n)(*[_as_rast_args(a) for a in args_list])
        self._check(self._lib.rast_draw_frames(self._h, arr, n, C.c_void_p(frames_ptr) if frames_ptr else None,
                                               C.c_void_p(depths_ptr) if depths_ptr else None, 1), "rast_draw_frames")

    def d2h_bytes(self):
        """Bytes of frame / depth data copied device -> host by this renderer so far (sparse copies count what they move)."""
        return int(self._lib.rast_d2h_bytes(self._h))

    def sync(self):
        self._check(self._lib.rast_sync(self._h), "rast_sync")

    def _band_rows(self, height):
        return self._band[1] - self._band[0] if getattr(self, "_band", None) else height

    # ---- auxiliary ----
    def set_keep_visibility(self, on):
        """Keep the last frame's visibility keys of every call (triangle_ids / stats need it; costs a clear in the next call)."""
        self._check(self._lib.rast_set_keep_visibility(self._h, int(bool(on))), "rast_set_keep_visibility")

    def triangle_ids(self, width, height):
        out = np.empty((height, width), np.uint32)
        self._check(self._lib.rast_read_triangle_ids(self._h, _ptr(out)), "rast_read_triangle_ids")
        return out

    def depth_to_u8(self, width, height):
        out = np.empty((height, width), np.uint8)
        self._check(self._lib.rast_depth_to_u8(self._h, _ptr(out)), "rast_depth_to_u8")
        return out

    def stats(self):
        s = RastStats()
        self._check(self._lib.rast_get_stats(self._h, C.byref(s)), "rast_get_stats")
        return dict(triangles=s.triangles, front_facing=s.front_facing, queued_chunks=s.queued_chunks, visible_pixels=s.visible_pixels)

    def last_schedule(self):
        """{'setup': ..., 'raster': ..., 'shade': ...}: the kernel flavour each pass of the most recent batch took."""
        txt = self._lib.rast_last_schedule(self._h).decode()
        return dict(kv.split("=", 1) for kv in txt.split(" ") if "=" in kv) if txt else {}

    def pin_host(self, array):
        """Page-lock a numpy output buffer the caller keeps alive (rast_host_register); unpin_host before freeing it."""
        self._check(self._lib.rast_host_register(C.c_void_p(array.ctypes.data), array.nbytes), "rast_host_register")

    def unpin_host(self, array):
        self._lib.rast_host_unregister(C.c_void_p(array.ctypes.data))

    def set_retained_outputs(self, on):
        """The caller promises that the host buffers of a draw still hold what the previous host-buffer draw of this renderer wrote
        (the reference's spin loop reuses its buffers): only the changed rectangles are rewritten.  Same bytes in the buffers."""
        self._check(self._lib.rast_set_retained_outputs(self._h, int(bool(on))), "rast_set_retained_outputs")

    def set_profiling(self, on):
        self._check(self._lib.rast_set_profiling(self._h, int(bool(on))), "rast_set_profiling")

    def pass_ms(self):
        ms = (C.c_float * len(RAST_PASS_NAMES))()
        self._check(self._lib.rast_get_pass_ms(self._h, ms), "rast_get_pass_ms")
        return dict(zip(RAST_PASS_NAMES, [float(x) for x in ms]))

    def launch_count(self):
        return int(self._lib.rast_launch_count(self._h))

    def selftest_division(self, n_samples=1 << 30, seed=1):
        """Quotients (out of ~n_samples) where the shared-reciprocal division differs from IEEE division: expected 0."""
        bad = C.c_uint64(0)
        self._check(self._lib.rast_selftest_division(self._h, int(n_samples), int(seed), C.byref(bad)), "rast_selftest_division")
        return int(bad.value)

    def light_trans_dirs(self):
        return np.array([list(self._lights[i].trans_dir) for i in range(self.n_lights)], np.float32)


def frame_matrices(args):
    """(modelview, camera, normal_matrix, view), each float32[16] column-major -- drawing.cpp:222-229."""
    a = _as_rast_args(args)
    out = [np.zeros(16, np.float32) for _ in range(4)]
    _lib.load().rast_frame_matrices(C.byref(a), *[_ptr(o) for o in out])
    return tuple(out)


def fnv1a64(a):
    """FNV-1a-64 of a numpy array's bytes as 16 hex digits -- the checksum format of tests/golden/*.json."""
    a = np.ascontiguousarray(a)
    return "%016x" % _lib.load().rast_fnv1a64(a.ctypes.data, a.nbytes)


def spin_angle(ry0, k, n_frames):
    return float(_lib.load().rast_spin_angle(float(ry0), int(k), int(n_frames)))


_cached = {}
CONTENT_KEY_MAX_BYTES = 64 << 20  # scenes up to this size are fingerprinted on every call; larger ones are re-uploaded unless scene_version is given


def _fingerprint(arrays):
    """Content fingerprint of the scene arrays (shape, dtype and every byte): rast_hash64 runs at memory speed (Suzanne with
    its 12 MB texture: about a millisecond; zlib.crc32 + adler32 took six)."""
    lib = _lib.load()
    out = []
    for a in arrays:
        b = np.ascontiguousarray(a)
        out.append((b.shape, b.dtype.str, int(lib.rast_hash64(b.ctypes.data if b.size else None, b.nbytes, len(out) + 1))))
    return tuple(out)


# Output buffers the drop-in has page-locked (rast_host_register): draws into pageable memory are several times slower.  The
# registry holds a reference to each array, so a registered buffer cannot be freed under the driver; the oldest entries are
# unregistered when more than PINNED_OUTPUTS_MAX distinct buffers have been seen.
_pinned_outputs = {}
PINNED_OUTPUTS_MAX = 4


def _pin_output(a):
    if a is None or a.nbytes < (64 << 10):
        return
    key = (a.ctypes.data, a.nbytes)
    if key in _pinned_outputs:
        return
    lib = _lib.load()
    while len(_pinned_outputs) >= PINNED_OUTPUTS_MAX:
        old_key = next(iter(_pinned_outputs))
        lib.rast_host_unregister(C.c_void_p(old_key[0]))
        del _pinned_outputs[old_key]
    if lib.rast_host_register(C.c_void_p(a.ctypes.data), a.nbytes) == 0:
        _pinned_outputs[key] = a


def unpin_outputs():
    """Unregister every output buffer draw_frame() has page-locked."""
    lib = _lib.load()
    for key in list(_pinned_outputs):
        lib.rast_host_unregister(C.c_void_p(key[0]))
    _pinned_outputs.clear()


def _material_arrays(materials):
    out = []
    for m in materials:
        kd = m.kd if isinstance(m, Material) else m["kd"]
        tex = m.texels if isinstance(m, Material) else m.get("texels")
        out.append(np.asarray(kd, np.float32))
        out.append(np.zeros(0, np.float32) if tex is None else np.asarray(tex, np.float32))
    return out


def invalidate():
    """Forget the scene cached by draw_frame(): the next call uploads again.  (Page-locked output buffers stay: unpin_outputs().)"""
    for r in _cached.values():
        r.close()
    _cached.clear()


def _check_out(name, a, dtype, shape):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.shape != shape or not a.flags.c_contiguous or not a.flags.writeable:
        raise RastError("draw_frame: %s must be a writeable C-contiguous numpy array of dtype %s and shape %s (got %s)"
                        % (name, np.dtype(dtype).name, shape, "%s %s" % (getattr(a, "dtype", type(a).__name__), getattr(a, "shape", ""))))


def draw_frame(model_vertices, faces, model_vertnormals, vertuvs, lights, materials, arguments, frame_buffer, depth_buffer, device=0, scene_version=None):
    """Same argument order and meaning as the reference's draw_frame (headers/drawing.h:16-18).

    frame_buffer: uint8 [3,H,W] (CImg planar), depth_buffer: float32 [H,W] (or None); both are overwritten with
    the finished frame (callers of the reference always pass cleared buffers, renderer.cpp:85-86).
    lights: float32 [L,10]; columns 7..9 (trans_dir) are written like Light::transform does.

    The reference reads its vectors on every call.  Here the scene stays on the GPU between calls as long as it is
    the same scene: by content (a checksum of every array, for scenes up to CONTENT_KEY_MAX_BYTES -- an edit in place
    is seen), or, for larger scenes, by the caller's `scene_version` (any hashable; bump it after changing the arrays).
    A large scene without a version is uploaded on every call, like the reference re-reads it.  invalidate() drops
    the cached scene."""
    a = _as_rast_args(arguments)
    _check_out("frame_buffer", frame_buffer, np.uint8, (3, a.image_height, a.image_width))
    if depth_buffer is not None:
        _check_out("depth_buffer", depth_buffer, np.float32, (a.image_height, a.image_width))
    arrays = [np.asarray(model_vertices), np.asarray(faces), np.asarray(model_vertnormals), np.asarray(vertuvs)] + _material_arrays(materials)
    if scene_version is not None:
        key = (device, "version", scene_version)
    elif sum(x.nbytes for x in arrays) <= CONTENT_KEY_MAX_BYTES:
        key = (device, "content", _fingerprint(arrays))
    else:
        key = None
    r = _cached.get(key) if key is not None else None
    if r is None:
        invalidate()
        r = Renderer(device)
        r.upload_mesh(model_vertices, faces, model_vertnormals, vertuvs)
        r.upload_materials(materials)
        _cached[key if key is not None else (device, "uncached")] = r
    r.set_lights(np.asarray(lights, np.float32)[:, :7])
    _pin_output(frame_buffer)
    _pin_output(depth_buffer)
    r.draw_frame(a, frame_buffer, depth_buffer, want_depth=depth_buffer is not None)
    if isinstance(lights, np.ndarray) and lights.ndim == 2 and lights.shape[1] >= 10:
        lights[:, 7:10] = r.light_trans_dirs()
    if key is None:
        invalidate()
