"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the reference's own
hot path compiled from /root/reference (oracle/_ref/libref.so).  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
NO_TRIANGLE = 0xFFFFFFFF


class OrcLight(C.Structure):
    _fields_ = [("direction", C.c_float * 3), ("intensity", C.c_float), ("colour", C.c_float * 3), ("trans_dir", C.c_float * 3)]


class OrcMaterial(C.Structure):
    _fields_ = [("kd", C.c_float * 3), ("has_texture", C.c_int32), ("tex_w", C.c_int32), ("tex_h", C.c_int32), ("texels", C.c_void_p)]


class OrcScene(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("n_positions", C.c_uint32),
                ("normals", C.c_void_p), ("n_normals", C.c_uint32),
                ("uvs", C.c_void_p), ("n_uvs", C.c_uint32),
                ("tris", C.c_void_p), ("n_tris", C.c_uint64),
                ("materials", C.POINTER(OrcMaterial)), ("n_materials", C.c_uint32)]


class OrcArgs(C.Structure):
    _fields_ = [("image_width", C.c_uint32), ("image_height", C.c_uint32), ("aspect_ratio", C.c_float), ("scale", C.c_float),
                ("displacement", C.c_float * 3), ("tait_bryan_angles", C.c_float * 3), ("wind_clockwise", C.c_int32), ("flat", C.c_int32)]


class OrcCounters(C.Structure):
    _fields_ = [("front_facing", C.c_uint64), ("bbox_tests", C.c_uint64), ("covered", C.c_uint64), ("depth_passes", C.c_uint64)]


def build_oracle(force=False):
    """Compile oracle/liboracle.so (plain C, gcc) if missing or stale."""
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("oracle.c", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src[0], "-lm", "-lpthread"])
    return so


def build_ref():
    """Compile oracle/_ref/libref.so from /root/reference when that tree exists (dev container only)."""
    so = os.path.join(ORACLE_DIR, "_ref", "libref.so")
    if not os.path.exists(so) and os.path.isdir(os.environ.get("REFERENCE_ROOT", "/root/reference")):
        subprocess.check_call([os.path.join(ORACLE_DIR, "build_ref.sh")])
    return so if os.path.exists(so) else None


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.orc_signed_area_2d.restype = C.c_float
        lib.orc_spin_angle.restype = C.c_float
        lib.orc_spin_angle.argtypes = [C.c_float, C.c_uint32, C.c_uint32]
        lib.orc_fnv1a64.restype = C.c_uint64
        lib.orc_fnv1a64.argtypes = [C.c_void_p, C.c_uint64]
        lib.orc_transformation_matrix.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_camera_matrix.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        lib.orc_depth_to_u8.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_normalize_texture.argtypes = [C.c_void_p, C.c_uint64]
        lib.orc_draw_frame.argtypes = [C.POINTER(OrcScene), C.c_void_p, C.c_uint32, C.POINTER(OrcArgs), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_uint32, C.c_uint32, C.POINTER(OrcCounters)]
        lib.orc_draw_frame_mt.argtypes = [C.POINTER(OrcScene), C.c_void_p, C.c_uint32, C.POINTER(OrcArgs), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.POINTER(OrcCounters)]
        _oracle = lib
    return _oracle


_ref = None


def ref():
    """The reference's own compiled hot path, or None when it cannot exist (GPU box, no /root/reference)."""
    global _ref
    if _ref is None:
        so = build_ref()
        if so is None:
            return None
        lib = C.CDLL(so)
        lib.ref_scene_create.restype = C.c_void_p
        lib.ref_scene_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.POINTER(C.c_char_p), C.c_uint32]
        lib.ref_scene_destroy.argtypes = [C.c_void_p]
        lib.ref_scene_draw.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.ref_depth_to_u8.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.ref_transformation_matrix.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_camera_matrix.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        lib.ref_normal_matrix.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_transform_direction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_raster_vertex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.ref_signed_area_2d.restype = C.c_float
        lib.ref_signed_area_2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_transform_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        lib.ref_shade.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.ref_material_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.ref_load_obj.restype = C.c_void_p
        lib.ref_load_obj.argtypes = [C.c_char_p, C.c_char_p]
        lib.ref_scene_sizes.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_scene_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_load_lights.restype = C.c_uint32
        lib.ref_load_lights.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32]
        _ref = lib
    return _ref


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Scene:
    """Flat arrays in the layout the reference's vectors hold (renderer.cpp:55-72).

    materials: list of dicts {kd: (r,g,b), texels: None | float32 [3,h,w] already normalised}."""

    def __init__(self, positions, normals, uvs, tris, materials):
        self.positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        self.normals = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        self.uvs = np.ascontiguousarray(uvs, np.float32).reshape(-1, 2)
        self.tris = np.ascontiguousarray(tris, np.int32).reshape(-1, 10)
        self.materials = materials
        for m in self.materials:
            if m.get("texels") is not None:
                m["texels"] = np.ascontiguousarray(m["texels"], np.float32)

    def orc(self):
        mats = (OrcMaterial * max(1, len(self.materials)))()
        for i, m in enumerate(self.materials):
            mats[i].kd = (C.c_float * 3)(*m["kd"])
            t = m.get("texels")
            mats[i].has_texture = 0 if t is None else (3 if m.get("modulate_kd") else 1)
            if t is not None:
                mats[i].tex_h, mats[i].tex_w = t.shape[1], t.shape[2]
                mats[i].texels = t.ctypes.data
        s = OrcScene(self.positions.ctypes.data, len(self.positions), self.normals.ctypes.data, len(self.normals),
                     self.uvs.ctypes.data, len(self.uvs), self.tris.ctypes.data, len(self.tris), mats, len(self.materials))
        s._keep = (mats, self)
        return s


def make_args(width, height, scale=1.0, disp=(0, 0, 0), angles=(0, 0, 0), wind_clockwise=False, flat=False):
    a = OrcArgs()
    a.image_width, a.image_height = width, height
    a.aspect_ratio = float(np.float32(width) / np.float32(height))  # arguments.cpp:39
    a.scale = scale
    a.displacement = (C.c_float * 3)(*disp)
    a.tait_bryan_angles = (C.c_float * 3)(*angles)
    a.wind_clockwise = int(wind_clockwise)
    a.flat = int(flat)
    return a


def lights_array(lights):
    """lights: array-like [L,7] (dx,dy,dz,intensity,r,g,b) -> float32 [L,10] with trans_dir zeroed."""
    l7 = np.asarray(lights, np.float32).reshape(-1, 7)
    out = np.zeros((len(l7), 10), np.float32)
    out[:, :7] = l7
    return out


def oracle_draw(scene, lights, args, threads=1, band=None, want_counters=False):
    """Run the oracle on cleared buffers.  Returns (frame u8 [3,H,W], depth f32 [H,W], tri_id u32 [H,W][, counters])."""
    lib = oracle()
    W, H = args.image_width, args.image_height
    frame = np.zeros((3, H, W), np.uint8)
    depth = np.ones((H, W), np.float32)
    tri = np.full((H, W), NO_TRIANGLE, np.uint32)
    l10 = lights_array(np.asarray(lights, np.float32).reshape(-1, 10)[:, :7] if np.asarray(lights).shape[-1] == 10 else lights)
    cnt = OrcCounters()
    s = scene.orc()
    if threads > 1 and band is None:
        lib.orc_draw_frame_mt(C.byref(s), ptr(l10), len(l10), C.byref(args), ptr(frame), ptr(depth), ptr(tri), threads, C.byref(cnt))
    else:
        y0, y1 = band if band is not None else (0, H)
        lib.orc_draw_frame(C.byref(s), ptr(l10), len(l10), C.byref(args), ptr(frame), ptr(depth), ptr(tri), y0, y1, C.byref(cnt))
    if want_counters:
        return frame, depth, tri, cnt
    return frame, depth, tri


def fnv(a):
    a = np.ascontiguousarray(a)
    return "%016x" % oracle().orc_fnv1a64(ptr(a), a.nbytes)


def load_texture_png(path):
    """PNG -> float32 [3,h,w], min/max normalised over all channels jointly like Material's ctor (material.h:20-23)."""
    from PIL import Image
    im = np.asarray(Image.open(path).convert("RGB"), np.uint8)
    t = np.ascontiguousarray(im.transpose(2, 0, 1).astype(np.float32))
    oracle().orc_normalize_texture(ptr(t), t.size)
    return t


def build_shim_real_headers(force=False):
    """oracle/_ref/shim_real_headers: tests/shim_real_headers_main.cpp compiled against the reference's REAL headers (and its
    unmodified fileloader.cpp) where they lie under /root/reference, with INTEGRATION.md section 3's one-line patch applied to a
    temporary copy of headers/material.h.  Dev container only; the binary travels to the GPU box with oracle/_ref/."""
    import shutil
    import tempfile
    ref_root = os.environ.get("REFERENCE_ROOT", "/root/reference")
    out = os.path.join(ORACLE_DIR, "_ref", "shim_real_headers")
    src = os.path.join(ROOT, "tests", "shim_real_headers_main.cpp")
    hdr = os.path.join(ROOT, "include", "rast_draw_frame.hpp")
    lib = os.path.join(ROOT, "rasteriser_b200", "librast_b200.so")
    if not os.path.isdir(ref_root):
        return out if os.path.exists(out) else None
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(f) for f in (src, hdr)):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="shim_hdrs_")
    try:
        for f in os.listdir(os.path.join(ref_root, "headers")):  # quote-includes resolve next to the including file: mirror the directory
            if f != "material.h":
                os.symlink(os.path.join(ref_root, "headers", f), os.path.join(tmp, f))
        text = open(os.path.join(ref_root, "headers", "material.h")).read()
        assert "public:" in text and "RastMaterialView" not in text
        open(os.path.join(tmp, "material.h"), "w").write(text.replace("public:", "public:\n\tfriend struct RastMaterialView;", 1))
        flags = ["-std=c++11", "-O2", "-ffp-contract=off", "-Dcimg_display=0", "-w", "-I" + tmp, "-I" + os.path.join(ORACLE_DIR, "glm_stub"),
                 "-I" + os.path.join(ref_root, "vendor", "cimg"), "-I" + os.path.join(ref_root, "vendor", "tinyobjloader"), "-I" + os.path.join(ROOT, "include")]
        subprocess.check_call(["g++"] + flags + ["-o", out, src, os.path.join(ref_root, "fileloader.cpp"), "-L" + os.path.dirname(lib), "-lrast_b200",
                                                 "-Wl,-rpath,$ORIGIN/../../rasteriser_b200", "-lpthread"])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out
