"""Synthetic meshes of the shapes BASELINE.json names, generated on the host
with numpy from explicit formulas and a portable RNG, so inputs are reproducible on any toolchain.

  tessellate(...)     configs 3 / 5: every triangle of a base mesh subdivided n x n (n = 91 -> 8.0 M
                      triangles, n = 227 -> 49.9 M from Suzanne's 968)
  overdraw_scene(...) config 4: large overlapping random triangles, depth complexity ~50 at 8K
  random_lights(...)  config 5: 64 directional lights
"""
import numpy as np


def _splitmix64(seed, n):
    """n uniform floats in [0,1) from splitmix64: (x >> 40) * 2^-24."""
    with np.errstate(over="ignore"):
        z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(seed)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def tessellate(positions, normals, uvs, tris, n):
    """Subdivide every triangle n x n.  Vertex (i,j) of a triangle (A,B,C) is a*A + b*B + c*C with
    b = i/n, c = j/n, a = 1 - b - c (fp32, in that order); normals and uvs are interpolated the same
    way; vertices are not shared between original triangles.  Triangle order: original order, then i
    major, j minor, "up" (p,q,r) before "down" (q,s,r) -- the winding of the original is kept.
    Returns (positions, normals, uvs, tris[T*n*n, 10])."""
    positions = np.asarray(positions, np.float32)
    normals = np.asarray(normals, np.float32)
    uvs = np.asarray(uvs, np.float32).reshape(-1, 2)
    tris = np.asarray(tris, np.int32)
    T = len(tris)
    ij = np.array([(i, j) for i in range(n + 1) for j in range(n + 1 - i)], np.int32)
    per = len(ij)  # (n+1)(n+2)/2
    b = (ij[:, 0].astype(np.float32) / np.float32(n)).astype(np.float32)
    c = (ij[:, 1].astype(np.float32) / np.float32(n)).astype(np.float32)
    a = (np.float32(1.0) - b - c).astype(np.float32)
    local = -np.ones((n + 1, n + 1), np.int64)
    local[ij[:, 0], ij[:, 1]] = np.arange(per)

    def lerp(attr, idx, width):
        ok = idx >= 0
        src = attr[np.where(ok, idx, 0)] if len(attr) else np.zeros((T, 3, width), np.float32)
        A, B, C = src[:, 0], src[:, 1], src[:, 2]
        out = (a[None, :, None] * A[:, None, :] + b[None, :, None] * B[:, None, :]) + c[None, :, None] * C[:, None, :]
        return out.astype(np.float32).reshape(T * per, width)

    pos_out = lerp(positions, tris[:, 0:3], 3)
    nrm_out = lerp(normals, tris[:, 3:6], 3)
    has_uv = (tris[:, 6:9] >= 0).all(axis=1)
    uv_out = lerp(uvs, tris[:, 6:9], 2)

    up, down = [], []
    for i in range(n):
        for j in range(n - i):
            p, q, r = local[i, j], local[i + 1, j], local[i, j + 1]
            up.append((i, j, p, q, r))
            if j < n - 1 - i:
                down.append((i, j, q, local[i + 1, j + 1], r))
    # interleave: for each (i,j) the up triangle then, if it exists, the down triangle
    order = sorted([(i, j, 0, x, y, z) for i, j, x, y, z in up] + [(i, j, 1, x, y, z) for i, j, x, y, z in down])
    loc = np.array([(x, y, z) for _, _, _, x, y, z in order], np.int64)  # [n*n, 3]
    base = (np.arange(T, dtype=np.int64) * per)[:, None, None]
    v = (base + loc[None, :, :]).reshape(-1, 3)
    out = np.empty((T * n * n, 10), np.int32)
    out[:, 0:3] = v
    out[:, 3:6] = v
    out[:, 6:9] = np.where(np.repeat(has_uv, n * n)[:, None], v, -1)
    out[:, 9] = np.repeat(tris[:, 9], n * n)
    return pos_out, nrm_out, uv_out, out


def overdraw_scene(n_tris, width, height, radius_px=80.0, seed=12345):
    """Config 4: n_tris triangles of circumradius radius_px pixels at uniformly random positions and
    depths inside the frustum, counter-clockwise on screen (front-facing), one shared normal (0,0,1),
    no uvs, material 0.  Model space = the reference's default pose (view translate (0,0,-3),
    perspective fovy 45 deg, drawing.cpp:224-229)."""
    u = _splitmix64(seed, n_tris * 7).reshape(n_tris, 7)
    aspect = np.float32(width) / np.float32(height)
    tan_half = np.float32(np.tan(np.float32(np.pi / 8)))
    ze = -(1.5 + 3.0 * u[:, 0])
    cx, cy = 2.0 * u[:, 1] - 1.0, 2.0 * u[:, 2] - 1.0
    rn = radius_px / (0.5 * height)
    phase = 2.0 * np.pi * u[:, 3]
    pos = np.empty((n_tris, 3, 3), np.float32)
    for k in range(3):
        ang = phase + k * (2.0 * np.pi / 3.0)
        nx = cx + rn * np.cos(ang) / aspect
        ny = cy + rn * np.sin(ang)
        z = ze + 0.05 * (u[:, 4 + k] - 0.5)
        pos[:, k, 0] = nx * (-z) * aspect * tan_half
        pos[:, k, 1] = ny * (-z) * tan_half
        pos[:, k, 2] = z + 3.0
    tris = np.empty((n_tris, 10), np.int32)
    idx = np.arange(n_tris * 3, dtype=np.int32).reshape(n_tris, 3)
    tris[:, 0:3] = idx
    tris[:, 3:6] = 0
    tris[:, 6:9] = -1
    tris[:, 9] = 0
    return pos.reshape(-1, 3), np.array([[0, 0, 1]], np.float32), np.zeros((0, 2), np.float32), tris


def random_lights(n, seed=12345):
    """Config 5: n directional lights, dir = (2u-1, 2u-1, -u-0.1), intensity 1024/n, grey colour u."""
    u = _splitmix64(seed ^ 0xABCDEF, n * 4).reshape(n, 4)
    l = np.empty((n, 7), np.float32)
    l[:, 0], l[:, 1], l[:, 2] = 2 * u[:, 0] - 1, 2 * u[:, 1] - 1, -u[:, 2] - 0.1
    l[:, 3] = 1024.0 / n
    l[:, 4:7] = u[:, 3:4]
    return l


def write_obj(path, positions, normals, uvs, tris, mtllib=None, usemtl=None, digits=9):
    """Write the arrays as a Wavefront .obj (v / vn / vt / f i/j/k lines, 1-based) -- the input of the loader
    benchmarks and tests.  %.{digits}g text of an fp32 value is not guaranteed to parse back to the same bits
    through tinyobjloader's float reader; what matters is that the product loader and the reference's read the
    SAME text identically.  Written in slices so an 8 M-triangle file does not need its text in memory at once."""
    positions = np.asarray(positions, np.float32).reshape(-1, 3)
    normals = np.asarray(normals, np.float32).reshape(-1, 3)
    uvs = np.asarray(uvs, np.float32).reshape(-1, 2)
    tris = np.asarray(tris, np.int32).reshape(-1, 10)
    fmt = "%." + str(digits) + "g"
    with open(path, "w") as f:
        f.write("# written by rasteriser_b200.synth.write_obj\n")
        if mtllib:
            f.write("mtllib %s\n" % mtllib)
        for tag, arr in (("v", positions), ("vn", normals), ("vt", uvs)):
            line = tag + " " + " ".join([fmt] * arr.shape[1]) + "\n"
            for s in range(0, len(arr), 1 << 18):
                f.write("".join(line % tuple(r) for r in arr[s:s + (1 << 18)].tolist()))
        if usemtl:
            f.write("usemtl %s\n" % usemtl)
        has_n, has_t = len(normals) > 0, len(uvs) > 0
        for s in range(0, len(tris), 1 << 18):
            t = tris[s:s + (1 << 18)].astype(np.int64) + 1
            if has_n and has_t:
                f.write("".join("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (r[0], r[6], r[3], r[1], r[7], r[4], r[2], r[8], r[5]) for r in t.tolist()))
            elif has_n:
                f.write("".join("f %d//%d %d//%d %d//%d\n" % (r[0], r[3], r[1], r[4], r[2], r[5]) for r in t.tolist()))
            else:
                f.write("".join("f %d %d %d\n" % (r[0], r[1], r[2]) for r in t.tolist()))
