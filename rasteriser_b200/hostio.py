"""ctypes access to librast_host.so: the product's host-side C++ loaders (OBJ / MTL / lights CSV / texture),
PNG codec and flag parser (rasteriser_b200/host/).  Used by bench.py and the tests; the renderer executable
links the same sources directly."""
import ctypes as C
import os

import numpy as np

from . import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build.build_host_lib())
        l.rasth_load_obj.restype = C.c_void_p
        l.rasth_load_obj.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        l.rasth_load_obj_mt.restype = C.c_void_p
        l.rasth_load_obj_mt.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_void_p, C.c_char_p, C.c_int]
        l.rasth_set_obj_piece_bytes.argtypes = [C.c_uint64]
        l.rasth_save_mesh_cache.argtypes = [C.c_void_p, C.c_char_p]
        l.rasth_load_mesh_cache.restype = C.c_void_p
        l.rasth_load_mesh_cache.argtypes = [C.c_char_p]
        l.rasth_model_free.argtypes = [C.c_void_p]
        l.rasth_model_sizes.argtypes = [C.c_void_p, C.c_void_p]
        l.rasth_model_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        l.rasth_model_material.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        l.rasth_load_lights.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        l.rasth_parse_float.restype = C.c_float
        l.rasth_parse_float.argtypes = [C.c_char_p]
        l.rasth_png_write.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        l.rasth_png_read.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64]
        l.rasth_parse_args.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        _lib = l
    return _lib


def load_obj(path, mats_dir="", threads=0, stats=None):
    """-> dict(pos, nrm, uv, tris, materials=[dict(kd, texels)]), warnings.  threads = OBJ parser threads (0 = all);
    stats: optional dict that receives the loader's timings."""
    l = lib()
    err = C.create_string_buffer(4096)
    st = (C.c_double * 7)()
    h = l.rasth_load_obj_mt(path.encode(), mats_dir.encode(), threads, st, err, 4096)
    if not h:
        raise RuntimeError(err.value.decode())
    if stats is not None:
        stats.update(dict(zip(("threads", "file_bytes", "read_s", "scan_s", "resolve_s", "parse_s", "total_s"), list(st))))
    return _model_to_dict(l, h), err.value.decode()


def load_mesh_cache(path):
    l = lib()
    h = l.rasth_load_mesh_cache(path.encode())
    if not h:
        raise RuntimeError("cannot read mesh cache " + path)
    return _model_to_dict(l, h)


def save_mesh_cache(obj_path, mats_dir, cache_path, threads=0):
    """Parse obj_path with the product loader and write its binary cache."""
    l = lib()
    err = C.create_string_buffer(4096)
    h = l.rasth_load_obj_mt(obj_path.encode(), mats_dir.encode(), threads, None, err, 4096)
    if not h:
        raise RuntimeError(err.value.decode())
    rc = l.rasth_save_mesh_cache(h, cache_path.encode())
    l.rasth_model_free(h)
    if rc != 0:
        raise RuntimeError("cannot write mesh cache " + cache_path)


def _model_to_dict(l, h):
    sz = np.zeros(5, np.uint64)
    l.rasth_model_sizes(h, sz.ctypes.data)
    pos, nrm = np.zeros((int(sz[0]), 3), np.float32), np.zeros((int(sz[1]), 3), np.float32)
    uv, tris = np.zeros((int(sz[2]), 2), np.float32), np.zeros((int(sz[3]), 10), np.int32)
    l.rasth_model_copy(h, pos.ctypes.data, nrm.ctypes.data, uv.ctypes.data, tris.ctypes.data)
    mats = []
    for i in range(int(sz[4])):
        kd, info = np.zeros(3, np.float32), np.zeros(3, np.int32)
        l.rasth_model_material(h, i, kd.ctypes.data, info.ctypes.data, None)
        tex = None
        if info[0]:
            tex = np.zeros((3, int(info[2]), int(info[1])), np.float32)
            l.rasth_model_material(h, i, kd.ctypes.data, info.ctypes.data, tex.ctypes.data)
        mats.append(dict(kd=tuple(float(x) for x in kd), texels=tex))
    l.rasth_model_free(h)
    return dict(pos=pos, nrm=nrm, uv=uv, tris=tris, materials=mats)


def load_lights(path):
    out = np.zeros((256, 7), np.float32)
    n = lib().rasth_load_lights(path.encode(), out.ctypes.data, 256)
    if n < 0:
        raise RuntimeError("cannot read " + path)
    return out[:n].copy()


def parse_args(argv):
    arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
    u, f = np.zeros(6, np.uint32), np.zeros(8, np.float32)
    s = C.create_string_buffer(8192)
    rc = lib().rasth_parse_args(len(argv), arr, u.ctypes.data, f.ctypes.data, s, 8192)
    obj, lights, mats, msg = (s.value.decode().split("\n", 3) + ["", "", "", ""])[:4]
    return rc, dict(width=int(u[0]), height=int(u[1]), spin=bool(u[2]), flat=bool(u[3]), wind_clockwise=bool(u[4]), frames=int(u[5]),
                    aspect=float(f[0]), scale=float(f[1]), disp=tuple(float(x) for x in f[2:5]), angles=tuple(float(x) for x in f[5:8]),
                    obj=obj, lights=lights, mats_dir=mats, message=msg)
