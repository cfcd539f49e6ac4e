"""ctypes binding of librast_b200.so (include/rast.h).  The library is the product; this module only
declares its symbols.  A missing library or a missing CUDA device is a hard error -- there is no
CPU path to fall back to."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# RAST_LIB selects another build of the same library (tools/build_variants.py: A/B runs of kernel variants)
LIB_PATH = os.environ.get("RAST_LIB") or os.path.join(PKG, "librast_b200.so")

RAST_PASS_NAMES = ("clear", "vertex", "setup", "raster", "shade")
NO_TRIANGLE = 0xFFFFFFFF


class RastLight(C.Structure):
    """rast_light -- headers/light.h:7-14"""
    _fields_ = [("direction", C.c_float * 3), ("intensity", C.c_float), ("colour", C.c_float * 3), ("trans_dir", C.c_float * 3)]


class RastMaterial(C.Structure):
    """rast_material -- headers/material.h:11-25"""
    _fields_ = [("kd", C.c_float * 3), ("has_texture", C.c_int32), ("tex_w", C.c_int32), ("tex_h", C.c_int32), ("texels", C.c_void_p)]


class RastArgs(C.Structure):
    """rast_args -- the fields of Args draw_frame consumes (headers/arguments.h:7-21)"""
    _fields_ = [("image_width", C.c_uint32), ("image_height", C.c_uint32), ("aspect_ratio", C.c_float), ("scale", C.c_float),
                ("displacement", C.c_float * 3), ("tait_bryan_angles", C.c_float * 3), ("wind_clockwise", C.c_int32), ("flat", C.c_int32)]


class RastStats(C.Structure):
    _fields_ = [("triangles", C.c_uint64), ("front_facing", C.c_uint64), ("queued_chunks", C.c_uint64), ("visible_pixels", C.c_uint64)]


# every symbol include/rast.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "rast_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rast_destroy": (None, [C.c_void_p]),
    "rast_last_error": (C.c_char_p, [C.c_void_p]),
    "rast_version": (C.c_char_p, []),
    "rast_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rast_use_own_stream": (C.c_int, [C.c_void_p]),
    "rast_upload_mesh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "rast_upload_materials": (C.c_int, [C.c_void_p, C.POINTER(RastMaterial), C.c_uint32]),
    "rast_set_lights": (C.c_int, [C.c_void_p, C.POINTER(RastLight), C.c_uint32]),
    "rast_frame_matrices": (None, [C.POINTER(RastArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rast_transform_lights": (None, [C.c_void_p, C.POINTER(RastLight), C.c_uint32]),
    "rast_spin_angle": (C.c_float, [C.c_float, C.c_uint32, C.c_uint32]),
    "rast_set_band": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "rast_set_output_plane_stride": (C.c_int, [C.c_void_p, C.c_uint64]),
    "rast_set_output_frame_stride": (C.c_int, [C.c_void_p, C.c_uint32]),
    "rast_fnv1a64": (C.c_uint64, [C.c_void_p, C.c_uint64]),
    "rast_device_alloc": (C.c_void_p, [C.c_void_p, C.c_uint64]),
    "rast_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rast_device_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "rast_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rast_ipc_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "rast_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rast_draw_frame": (C.c_int, [C.c_void_p, C.POINTER(RastArgs), C.c_void_p, C.c_void_p, C.POINTER(RastLight)]),
    "rast_draw_frame_device": (C.c_int, [C.c_void_p, C.POINTER(RastArgs), C.c_void_p, C.c_void_p]),
    "rast_draw_frames": (C.c_int, [C.c_void_p, C.POINTER(RastArgs), C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]),
    "rast_sync": (C.c_int, [C.c_void_p]),
    "rast_set_keep_visibility": (C.c_int, [C.c_void_p, C.c_int]),
    "rast_read_triangle_ids": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rast_depth_to_u8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rast_get_stats": (C.c_int, [C.c_void_p, C.POINTER(RastStats)]),
    "rast_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "rast_set_retained_outputs": (C.c_int, [C.c_void_p, C.c_int]),
    "rast_last_schedule": (C.c_char_p, [C.c_void_p]),
    "rast_host_register": (C.c_int, [C.c_void_p, C.c_uint64]),
    "rast_host_unregister": (C.c_int, [C.c_void_p]),
    "rast_hash64": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "rast_get_pass_ms": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rast_launch_count": (C.c_uint64, [C.c_void_p]),
    "rast_d2h_bytes": (C.c_uint64, [C.c_void_p]),
    "rast_selftest_division": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "rast_host_alloc": (C.c_void_p, [C.c_uint64]),
    "rast_host_free": (None, [C.c_void_p]),
}

_lib = None


def load():
    """Load librast_b200.so and declare its prototypes; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("librast_b200.so is not built (run `python -m rasteriser_b200.build`); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib
