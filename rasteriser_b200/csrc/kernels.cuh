// kernels.cuh -- sm_100a kernels of the frame path (vertex -> setup/cull -> rasterise with a 64-bit
// visibility buffer -> deferred resolve + shade).  Semantics follow the reference line by line
// (citations per kernel); the structure does not: the reference walks triangles one after another
// and shades every depth-passing fragment immediately (drawing.cpp:250-257, :119-146), this
// pipeline resolves visibility order-independently with atomicMin on (depth key << 32 | triangle
// index) and shades each pixel once.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "exact.cuh"

namespace rk {

constexpr unsigned long long VIS_EMPTY = ~0ull;
constexpr uint32_t INVALID_TRI = 0xFFFFFFFFu;
constexpr int CHUNK = 32;                   // a queued work item covers <= CHUNK x CHUNK pixels of a triangle's bbox
constexpr uint32_t TINY_MAX_PIXELS = 16;    // bboxes up to this many pixels are rasterised by the setup thread itself
constexpr float EDGE_SLACK = 2.384185791015625e-07f; // 2^-22, see candidate()

// ---- device-side data ----------------------------------------------------------------------
struct FrameParams {          // one per frame of a batch, built on the host (hostmath.h)
    float camera[16];         // perspective * view * model   (drawing.cpp:229)
    float normal_m[16];       // transpose(inverse(modelview)) (geometry.cpp:101)
    uint32_t wind_clockwise;  // arguments.h:14
    uint32_t pad[3];
};

struct LightDev {             // pre-combined per light: -trans_dir and intensity*colour (shading.cpp:21)
    float ntx, nty, ntz, icr, icg, icb, pad0, pad1;
};

struct MaterialDev {          // material.h:11-25
    float kd[3];
    int has_texture;
    int tex_w, tex_h;
    long long texel_offset;   // into Scene::texels, planar [3][h][w]
};

struct Scene {
    const float *pos;         // xyz [V]
    const float *nrm;         // xyz [Nn]
    const float2 *uv;         // [Nuv]
    const int *vidx0, *vidx1, *vidx2; // vertex indices, SoA [T] (coalesced in the per-triangle pass)
    const int4 *attr;         // [2T]: (n0,n1,n2,material), (t0,t1,t2,-) -- touched only for visible triangles
    const MaterialDev *mats;
    const float *texels;
    uint32_t V, Nn, Nuv, M;
    uint64_t T;
};

struct View {
    uint32_t W, H;            // image size (arguments.h:8-9)
    uint32_t y0, y1;          // band of rows rendered by this context, [y0,y1)
    uint32_t band_pixels;     // W * (y1 - y0)
};

struct Batch {
    const FrameParams *frames;
    uint32_t n_frames;
    float4 *rv;               // raster vertices (x, y, ndc z, 1/w) [n_frames][V]  (drawing.cpp:216,247)
    unsigned long long *vis;  // visibility buffer [n_frames][band_pixels]
    uint2 *queue;             // work items (triangle, cx | cy<<12 | frame<<24)
    uint32_t queue_cap;
    unsigned long long *counters; // [0] queue count (may exceed queue_cap), [1] queue cursor, [2] overflow flag
};

// ---- visibility key ------------------------------------------------------------------------
// Order-preserving map of a finite float below 1.0 to u32.  -0 is canonicalised so that it ties
// with +0 (the reference's strict '<' treats them as equal, drawing.cpp:119).
__device__ __forceinline__ uint32_t depth_key(float z) {
    uint32_t b = __float_as_uint(z);
    if (z == 0.f) b = 0u;
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// ---- triangle setup ------------------------------------------------------------------------
struct TriSetup {
    float x0, y0, x1, y1, x2, y2;
    float z0, z1, z2;
    float d12x, d12y, d20x, d20y, d01x, d01y; // the pixel-invariant differences of edge() (drawing.cpp:38)
    float area;                               // edge(v2; v0, v1) (drawing.cpp:46)
    uint32_t flip;                            // sign bit of area
    bool literal;                             // area is 0, inf or NaN: no sign shortcut
};

__device__ __forceinline__ void tri_setup(TriSetup &s, const float4 &v0, const float4 &v1, const float4 &v2) {
    s.x0 = v0.x; s.y0 = v0.y; s.z0 = v0.z;
    s.x1 = v1.x; s.y1 = v1.y; s.z1 = v1.z;
    s.x2 = v2.x; s.y2 = v2.y; s.z2 = v2.z;
    s.d12x = exact::sub(v2.x, v1.x); s.d12y = exact::sub(v2.y, v1.y);
    s.d20x = exact::sub(v0.x, v2.x); s.d20y = exact::sub(v0.y, v2.y);
    s.d01x = exact::sub(v1.x, v0.x); s.d01y = exact::sub(v1.y, v0.y);
    s.area = exact::sub(exact::mul(s.d01x, exact::sub(v2.y, v0.y)), exact::mul(s.d01y, exact::sub(v2.x, v0.x)));
    s.flip = __float_as_uint(s.area) & 0x80000000u;
    s.literal = !(fabsf(s.area) > 0.f && fabsf(s.area) < __int_as_float(0x7f800000));
}

// signed_area_2d (geometry.cpp:76-83), left to right
__device__ __forceinline__ float signed_area_2d(const float4 &v0, const float4 &v1, const float4 &v2) {
    float a = exact::sub(exact::mul(v0.x, v1.y), exact::mul(v1.x, v0.y));
    a = exact::add(a, exact::mul(v1.x, v2.y));
    a = exact::sub(a, exact::mul(v2.x, v1.y));
    a = exact::add(a, exact::mul(v2.x, v0.y));
    a = exact::sub(a, exact::mul(v0.x, v2.y));
    return exact::mul(-0.5f, a);
}

struct BBox { uint32_t x0, y0, x1, y1; bool empty; };

// bounding_box (drawing.cpp:77-93) intersected with the band [vy0, vy1)
__device__ __forceinline__ BBox bounding_box(const float4 &v0, const float4 &v1, const float4 &v2, const View &vw) {
    using namespace exact;
    const float brx = (float)(vw.W - 1u), bry = (float)(vw.H - 1u);
    const float minx = glm_min(glm_min(v0.x, v1.x), v2.x), miny = glm_min(glm_min(v0.y, v1.y), v2.y);
    const float maxx = ceilf(glm_max(glm_max(v0.x, v1.x), v2.x)), maxy = ceilf(glm_max(glm_max(v0.y, v1.y), v2.y));
    BBox b;
    b.x0 = to_uint(glm_min(glm_max(minx, 0.f), brx));
    b.y0 = to_uint(glm_min(glm_max(miny, 0.f), bry));
    b.x1 = to_uint(glm_min(glm_max(maxx, 0.f), brx));
    b.y1 = to_uint(glm_min(glm_max(maxy, 0.f), bry));
    if (b.y0 < vw.y0) b.y0 = vw.y0;
    if (b.y1 >= vw.y1) b.y1 = vw.y1 - 1u;
    b.empty = (b.x1 < b.x0) || (b.y1 < b.y0);
    return b;
}

// The three edge functions of one pixel, each a fresh evaluation in the reference's order
// (drawing.cpp:36-39); incremental stepping would round differently.
__device__ __forceinline__ void edges(const TriSetup &s, float px, float py, float &e0, float &e1, float &e2) {
    using namespace exact;
    e0 = sub(mul(s.d12x, sub(py, s.y1)), mul(s.d12y, sub(px, s.x1)));
    e1 = sub(mul(s.d20x, sub(py, s.y2)), mul(s.d20y, sub(px, s.x2)));
    e2 = sub(mul(s.d01x, sub(py, s.y0)), mul(s.d01y, sub(px, s.x0)));
}

// Cheap superset of the inside test.  The reference tests e_i/area >= 0 (drawing.cpp:46-48,111).
// For finite non-zero area the quotient is >= 0 (counting -0) iff e_i has area's sign, is zero, or
// the quotient underflows to -0; underflow needs |e_i| <= 2^-150 * |area| < 2^-22.  So
// "(e_i with area's sign folded in) >= -2^-22 for all i" never rejects a pixel the exact test
// accepts; survivors take the literal divisions (needed for depth anyway).
__device__ __forceinline__ bool candidate(const TriSetup &s, float e0, float e1, float e2) {
    if (s.literal) return true;
    const float t0 = __uint_as_float(__float_as_uint(e0) ^ s.flip);
    const float t1 = __uint_as_float(__float_as_uint(e1) ^ s.flip);
    const float t2 = __uint_as_float(__float_as_uint(e2) ^ s.flip);
    return (t0 >= -EDGE_SLACK) && (t1 >= -EDGE_SLACK) && (t2 >= -EDGE_SLACK);
}

// barycentric + inside + depth (drawing.cpp:41-49,111,115-119).  True iff the fragment is inside
// and nearer than the cleared depth 1.0f (a fragment at z >= 1 or NaN can never pass the strict '<').
__device__ __forceinline__ bool fragment(const TriSetup &s, float e0, float e1, float e2, float &b0, float &b1, float &b2, float &z) {
    using namespace exact;
    b0 = div(e0, s.area);
    b1 = div(e1, s.area);
    b2 = div(e2, s.area);
    if (!(b0 >= 0.f && b1 >= 0.f && b2 >= 0.f)) return false;
    z = add(add(mul(s.z0, b0), mul(s.z1, b1)), mul(s.z2, b2));
    return z < 1.0f;
}

__device__ __forceinline__ void test_and_commit(const TriSetup &s, uint32_t x, uint32_t y, uint32_t tri, unsigned long long *vis_row0, const View &vw) {
    float e0, e1, e2, b0, b1, b2, z;
    edges(s, (float)x, (float)y, e0, e1, e2);
    if (!candidate(s, e0, e1, e2)) return;
    if (!fragment(s, e0, e1, e2, b0, b1, b2, z)) return;
    const unsigned long long key = ((unsigned long long)depth_key(z) << 32) | tri;
    atomicMin(vis_row0 + (size_t)(y - vw.y0) * vw.W + x, key);
}

// ---- K0: clear ------------------------------------------------------------------------------
// renderer.cpp:85-86 / :107-108 (frame = 0, depth = 1.0f) become "no triangle" in the visibility buffer.
__global__ void k_clear(unsigned long long *vis, size_t n, unsigned long long *counters) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (blockIdx.x == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0ull;
    if (i + 1 < n) {
        *reinterpret_cast<ulonglong2 *>(vis + i) = make_ulonglong2(VIS_EMPTY, VIS_EMPTY);
    } else if (i < n) {
        vis[i] = VIS_EMPTY;
    }
}

// ---- K1: vertex stage -----------------------------------------------------------------------
// transform_point + z_divide + ndc_to_raster (geometry.cpp:44-74, drawing.cpp:241-247) fused: one
// thread per vertex, 12 B in, one float4 out.
__global__ void __launch_bounds__(256) k_vertex(Scene sc, View vw, Batch bt) {
    __shared__ float cam[16];
    const uint32_t f = blockIdx.y;
    if (threadIdx.x < 16) cam[threadIdx.x] = bt.frames[f].camera[threadIdx.x];
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sc.V) return;
    using namespace exact;
    const float x = sc.pos[3 * (size_t)i], y = sc.pos[3 * (size_t)i + 1], z = sc.pos[3 * (size_t)i + 2];
    const float4 clip = mat_vec(cam, x, y, z, 1.f);
    const float nx = div(clip.x, clip.w), ny = div(clip.y, clip.w), nz = div(clip.z, clip.w), nw = div(1.f, clip.w);
    float4 r;
    r.x = mul(mul(0.5f, add(nx, 1.0f)), (float)(int)vw.W);   // remap_ndc(x, width)
    r.y = mul(mul(0.5f, add(-ny, 1.0f)), (float)(int)vw.H);  // remap_ndc(-y, height)
    r.z = nz;
    r.w = nw;
    bt.rv[(size_t)f * sc.V + i] = r;
}

// ---- K2: triangle setup, cull, classification -----------------------------------------------
// draw_triangle up to the pixel loops (drawing.cpp:165-188).  Tiny bboxes are rasterised here;
// larger ones are cut into CHUNK x CHUNK work items for k_raster_chunks.
__global__ void __launch_bounds__(256) k_setup(Scene sc, View vw, Batch bt) {
    const uint32_t f = blockIdx.y;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= sc.T) return;
    const float4 *rv = bt.rv + (size_t)f * sc.V;
    const float4 v0 = rv[sc.vidx0[t]], v1 = rv[sc.vidx1[t]], v2 = rv[sc.vidx2[t]];

    const bool cw = bt.frames[f].wind_clockwise != 0u;
    const float a2 = signed_area_2d(v0, v1, v2);
    if (!((a2 > 0.f) != cw)) return; // back face (drawing.cpp:178-180)

    const BBox bb = bounding_box(v0, v1, v2, vw);
    if (bb.empty) return;
    const uint32_t w = bb.x1 - bb.x0 + 1u, h = bb.y1 - bb.y0 + 1u;
    unsigned long long *vis = bt.vis + (size_t)f * vw.band_pixels;

    bool inline_raster = (uint64_t)w * h <= TINY_MAX_PIXELS;
    if (!inline_raster) {
        const uint32_t ncx = (w + CHUNK - 1) / CHUNK, ncy = (h + CHUNK - 1) / CHUNK;
        const uint64_t n = (uint64_t)ncx * ncy;
        // reserve n consecutive slots; the 64-bit count keeps growing past the capacity, so it cannot wrap
        const uint64_t first = (n <= bt.queue_cap) ? atomicAdd(&bt.counters[0], (unsigned long long)n) : (uint64_t)bt.queue_cap;
        if (first + n > bt.queue_cap) {
            // queue full: void any slots reserved below the capacity and walk the whole bbox in this
            // thread (correct, slow); the host sees the flag and grows the queue for later frames
            for (uint64_t k = first; k < bt.queue_cap; ++k) bt.queue[k] = make_uint2(INVALID_TRI, 0u);
            bt.counters[2] = 1ull;
            inline_raster = true;
        } else {
            uint64_t k = first;
            for (uint32_t cy = 0; cy < ncy; ++cy)
                for (uint32_t cx = 0; cx < ncx; ++cx) bt.queue[k++] = make_uint2((uint32_t)t, cx | (cy << 12) | (f << 24));
        }
    }
    if (inline_raster) {
        TriSetup s;
        tri_setup(s, v0, v1, v2);
        for (uint32_t y = bb.y0; y <= bb.y1; ++y)
            for (uint32_t x = bb.x0; x <= bb.x1; ++x) test_and_commit(s, x, y, (uint32_t)t, vis, vw);
    }
}

// ---- K3: chunk rasteriser -------------------------------------------------------------------
// One warp per work item; lanes own 2x2 pixel quads of a 16x8 block, so the per-pixel
// differences (p - a) and half of the products of edge() are shared inside the quad.  A ballot
// skips blocks no lane may cover.  update_pixel's coverage + depth part (drawing.cpp:108-121).
__global__ void __launch_bounds__(256) k_raster_chunks(Scene sc, View vw, Batch bt) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t count = (uint32_t)min(bt.counters[0], (unsigned long long)bt.queue_cap);
    const uint32_t qx = (lane & 7u) * 2u, qy = (lane >> 3) * 2u;
    for (;;) {
        uint32_t e = 0;
        if (lane == 0) e = (uint32_t)min(atomicAdd(&bt.counters[1], 1ull), 0xFFFFFFFFull);
        e = __shfl_sync(0xFFFFFFFFu, e, 0);
        if (e >= count) break;
        const uint2 item = bt.queue[e];
        const uint32_t tri = item.x;
        if (tri == INVALID_TRI) continue;
        const uint32_t cx = item.y & 0xFFFu, cy = (item.y >> 12) & 0xFFFu, f = item.y >> 24;
        const float4 *rv = bt.rv + (size_t)f * sc.V;
        const float4 v0 = rv[sc.vidx0[tri]], v1 = rv[sc.vidx1[tri]], v2 = rv[sc.vidx2[tri]];
        const BBox bb = bounding_box(v0, v1, v2, vw);
        TriSetup s;
        tri_setup(s, v0, v1, v2);
        unsigned long long *vis = bt.vis + (size_t)f * vw.band_pixels;

        const uint32_t rx0 = bb.x0 + cx * CHUNK, ry0 = bb.y0 + cy * CHUNK;
        const uint32_t rx1 = min(bb.x1, rx0 + CHUNK - 1u), ry1 = min(bb.y1, ry0 + CHUNK - 1u);
        for (uint32_t by = ry0; by <= ry1; by += 8u) {
            for (uint32_t bx = rx0; bx <= rx1; bx += 16u) {
                using namespace exact;
                const uint32_t x = bx + qx, y = by + qy;
                const float pxa = (float)x, pxb = (float)(x + 1u), pya = (float)y, pyb = (float)(y + 1u);
                // edge k at pixel (i,j): mul(dkx, py_j - yk) - mul(dky, px_i - xk)
                const float a0a = mul(s.d12x, sub(pya, s.y1)), a0b = mul(s.d12x, sub(pyb, s.y1));
                const float a1a = mul(s.d20x, sub(pya, s.y2)), a1b = mul(s.d20x, sub(pyb, s.y2));
                const float a2a = mul(s.d01x, sub(pya, s.y0)), a2b = mul(s.d01x, sub(pyb, s.y0));
                const float c0a = mul(s.d12y, sub(pxa, s.x1)), c0b = mul(s.d12y, sub(pxb, s.x1));
                const float c1a = mul(s.d20y, sub(pxa, s.x2)), c1b = mul(s.d20y, sub(pxb, s.x2));
                const float c2a = mul(s.d01y, sub(pxa, s.x0)), c2b = mul(s.d01y, sub(pxb, s.x0));
                float e0[4], e1[4], e2[4]; // (xa,ya) (xb,ya) (xa,yb) (xb,yb)
                e0[0] = sub(a0a, c0a); e0[1] = sub(a0a, c0b); e0[2] = sub(a0b, c0a); e0[3] = sub(a0b, c0b);
                e1[0] = sub(a1a, c1a); e1[1] = sub(a1a, c1b); e1[2] = sub(a1b, c1a); e1[3] = sub(a1b, c1b);
                e2[0] = sub(a2a, c2a); e2[1] = sub(a2a, c2b); e2[2] = sub(a2b, c2a); e2[3] = sub(a2b, c2b);
                uint32_t mask = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool in_rect = (x + (k & 1)) <= rx1 && (y + (k >> 1)) <= ry1;
                    if (in_rect && candidate(s, e0[k], e1[k], e2[k])) mask |= 1u << k;
                }
                if (!__any_sync(0xFFFFFFFFu, mask != 0u)) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (mask & (1u << k)) {
                        float b0, b1, b2, z;
                        if (fragment(s, e0[k], e1[k], e2[k], b0, b1, b2, z)) {
                            const unsigned long long key = ((unsigned long long)depth_key(z) << 32) | tri;
                            atomicMin(vis + (size_t)(y + (k >> 1) - vw.y0) * vw.W + (x + (k & 1)), key);
                        }
                    }
                }
            }
        }
    }
}

// ---- K4: resolve + deferred shading ---------------------------------------------------------
// CImg::_linear_atXY (CImg.h:13475-13492) on one channel plane
__device__ __forceinline__ float linear_at(const float *plane, int w, int h, float fx, float fy) {
    using namespace exact;
    const float hx = (float)(w - 1), hy = (float)(h - 1);
    const float nfx = fx < 0.f ? 0.f : (fx > hx ? hx : fx); // cimg::cut (CImg.h:5184-5186)
    const float nfy = fy < 0.f ? 0.f : (fy > hy ? hy : fy);
    const uint32_t x = to_uint(nfx), y = to_uint(nfy);
    const float dx = sub(nfx, (float)x), dy = sub(nfy, (float)y);
    const uint32_t nx = dx > 0.f ? x + 1u : x, ny = dy > 0.f ? y + 1u : y;
    const float Icc = plane[x + (size_t)y * w], Inc = plane[nx + (size_t)y * w];
    const float Icn = plane[x + (size_t)ny * w], Inn = plane[nx + (size_t)ny * w];
    const float t1 = sub(sub(add(Icc, Inn), Icn), Inc);
    const float t2 = add(sub(Inc, Icc), mul(dy, t1));
    return add(add(Icc, mul(dx, t2)), mul(dy, sub(Icn, Icc)));
}

struct Shaded { uint32_t r, g, b; float depth; };

// The shading half of update_pixel (drawing.cpp:121-146) for the winning triangle of one pixel.
__device__ __forceinline__ Shaded shade_pixel(unsigned long long key, uint32_t x, uint32_t y, const Scene &sc, const float4 *rv,
                                              const FrameParams &fp, const LightDev *lights, uint32_t n_lights) {
    using namespace exact;
    Shaded out;
    if (key == VIS_EMPTY) { out.r = out.g = out.b = 0u; out.depth = 1.0f; return out; }
    const uint32_t tri = (uint32_t)key;
    const float4 v0 = rv[sc.vidx0[tri]], v1 = rv[sc.vidx1[tri]], v2 = rv[sc.vidx2[tri]];
    TriSetup s;
    tri_setup(s, v0, v1, v2);
    float e0, e1, e2;
    edges(s, (float)x, (float)y, e0, e1, e2);
    const float b0 = div(e0, s.area), b1 = div(e1, s.area), b2 = div(e2, s.area);
    out.depth = add(add(mul(v0.z, b0), mul(v1.z, b1)), mul(v2.z, b2)); // same bits as the raster pass

    // interpolation_coords + camera-space depth (drawing.cpp:125-128)
    const float i0 = mul(v0.w, b0), i1 = mul(v1.w, b1), i2 = mul(v2.w, b2);
    const float d = div(1.f, add(add(i0, i1), i2));

    const int4 an = sc.attr[2 * (size_t)tri], at = sc.attr[2 * (size_t)tri + 1];
    // transform_direction on the three vertex normals (geometry.cpp:35-42,97-108), done lazily here
    float3 n[3];
    const int nidx[3] = {an.x, an.y, an.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float mx = 0.f, my = 0.f, mz = 0.f;
        if (nidx[k] >= 0) { mx = sc.nrm[3 * (size_t)nidx[k]]; my = sc.nrm[3 * (size_t)nidx[k] + 1]; mz = sc.nrm[3 * (size_t)nidx[k] + 2]; }
        const float4 t = mat_vec(fp.normal_m, mx, my, mz, 0.f);
        n[k] = make_float3(t.x, t.y, t.z);
    }
    // perspective_interpolate + normalize (drawing.cpp:64-75,131-132)
    const float mx = mul(d, add(add(mul(i0, n[0].x), mul(i1, n[1].x)), mul(i2, n[2].x)));
    const float my = mul(d, add(add(mul(i0, n[0].y), mul(i1, n[1].y)), mul(i2, n[2].y)));
    const float mz = mul(d, add(add(mul(i0, n[0].z), mul(i1, n[1].z)), mul(i2, n[2].z)));
    const float inv = div(1.f, fsqrt(add(add(mul(mx, mx), mul(my, my)), mul(mz, mz))));
    float nx = mul(mx, inv), ny = mul(my, inv), nz = mul(mz, inv);
    if (fp.wind_clockwise) { nx = -nx; ny = -ny; nz = -nz; }

    // Material::sample (material.cpp:11-26); material -1 = untextured white (unpinned corner)
    float ar = 1.f, ag = 1.f, ab = 1.f;
    if (an.w >= 0 && (uint32_t)an.w < sc.M) {
        const MaterialDev m = sc.mats[an.w];
        if (m.has_texture) {
            float2 uv[3];
            const int tidx[3] = {at.x, at.y, at.z};
#pragma unroll
            for (int k = 0; k < 3; ++k) uv[k] = tidx[k] >= 0 ? sc.uv[tidx[k]] : make_float2(0.f, 0.f);
            const float u = mul(d, add(add(mul(i0, uv[0].x), mul(i1, uv[1].x)), mul(i2, uv[2].x))); // drawing.cpp:135
            const float v = mul(d, add(add(mul(i0, uv[0].y), mul(i1, uv[1].y)), mul(i2, uv[2].y)));
            const float fx = mul(u, (float)m.tex_w), fy = mul(sub(1.f, v), (float)m.tex_h);
            const float *tex = sc.texels + m.texel_offset;
            const size_t plane = (size_t)m.tex_w * m.tex_h;
            ar = linear_at(tex, m.tex_w, m.tex_h, fx, fy);
            ag = linear_at(tex + plane, m.tex_w, m.tex_h, fx, fy);
            ab = linear_at(tex + 2 * plane, m.tex_w, m.tex_h, fx, fy);
        } else {
            ar = m.kd[0]; ag = m.kd[1]; ab = m.kd[2];
        }
    }

    // shade / light_contribution (shading.cpp:20-34)
    float sr = 0.f, sg = 0.f, sb = 0.f;
    for (uint32_t l = 0; l < n_lights; ++l) {
        const LightDev L = lights[l];
        const float k = glm_max(0.f, add(add(mul(nx, L.ntx), mul(ny, L.nty)), mul(nz, L.ntz)));
        sr = add(sr, mul(mul(mul(L.icr, ar), k), 0.318309886183790671537767526745028724f));
        sg = add(sg, mul(mul(mul(L.icg, ag), k), 0.318309886183790671537767526745028724f));
        sb = add(sb, mul(mul(mul(L.icb, ab), k), 0.318309886183790671537767526745028724f));
    }
    out.r = to_uint(glm_min(sr, 255.f)) & 0xFFu;
    out.g = to_uint(glm_min(sg, 255.f)) & 0xFFu;
    out.b = to_uint(glm_min(sb, 255.f)) & 0xFFu;
    return out;
}

// One thread per 4 consecutive pixels: keys in as 2 x 16 B, colour out as one uchar4 per plane
// and depth as one float4 (CImg planar layout, CImg.h:11715-11721).  VEC requires
// band_pixels % 4 == 0 and 16-byte aligned outputs; otherwise the scalar variant runs.
template <bool VEC>
__global__ void __launch_bounds__(256) k_resolve_shade(Scene sc, View vw, Batch bt, const LightDev *lights, uint32_t n_lights,
                                                       uint8_t *rgb, float *depth) {
    __shared__ FrameParams fp;
    __shared__ LightDev sl[64];
    const uint32_t f = blockIdx.y;
    for (uint32_t i = threadIdx.x; i < sizeof(FrameParams) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(&fp)[i] = reinterpret_cast<const uint32_t *>(&bt.frames[f])[i];
    const uint32_t nl_s = n_lights <= 64u ? n_lights : 0u;
    for (uint32_t i = threadIdx.x; i < nl_s * (sizeof(LightDev) / 4); i += blockDim.x) reinterpret_cast<uint32_t *>(sl)[i] = reinterpret_cast<const uint32_t *>(lights)[i];
    __syncthreads();
    const LightDev *lp = nl_s ? sl : lights;

    const uint32_t P = vw.band_pixels;
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= P) return;
    const unsigned long long *vis = bt.vis + (size_t)f * P;
    const float4 *rv = bt.rv + (size_t)f * sc.V;
    uint8_t *out_r = rgb + (size_t)f * 3 * P, *out_g = out_r + P, *out_b = out_g + P;
    float *out_d = depth ? depth + (size_t)f * P : nullptr;

    unsigned long long keys[4];
    if (VEC) {
        const ulonglong2 k01 = *reinterpret_cast<const ulonglong2 *>(vis + i4), k23 = *reinterpret_cast<const ulonglong2 *>(vis + i4 + 2);
        keys[0] = k01.x; keys[1] = k01.y; keys[2] = k23.x; keys[3] = k23.y;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) keys[k] = (i4 + k < P) ? vis[i4 + k] : VIS_EMPTY;
    }
    Shaded px[4];
    uint32_t x = (uint32_t)(i4 % vw.W), y = vw.y0 + (uint32_t)(i4 / vw.W);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        px[k] = shade_pixel(keys[k], x, y, sc, rv, fp, lp, n_lights);
        if (++x == vw.W) { x = 0; ++y; }
    }
    if (VEC) {
        *reinterpret_cast<uchar4 *>(out_r + i4) = make_uchar4(px[0].r, px[1].r, px[2].r, px[3].r);
        *reinterpret_cast<uchar4 *>(out_g + i4) = make_uchar4(px[0].g, px[1].g, px[2].g, px[3].g);
        *reinterpret_cast<uchar4 *>(out_b + i4) = make_uchar4(px[0].b, px[1].b, px[2].b, px[3].b);
        if (out_d) *reinterpret_cast<float4 *>(out_d + i4) = make_float4(px[0].depth, px[1].depth, px[2].depth, px[3].depth);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i4 + k < P) {
                out_r[i4 + k] = (uint8_t)px[k].r; out_g[i4 + k] = (uint8_t)px[k].g; out_b[i4 + k] = (uint8_t)px[k].b;
                if (out_d) out_d[i4 + k] = px[k].depth;
            }
        }
    }
}

// ---- auxiliary kernels ----------------------------------------------------------------------
__global__ void k_extract_tri_ids(const unsigned long long *vis, uint32_t *ids, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = (vis[i] == VIS_EMPTY) ? INVALID_TRI : (uint32_t)vis[i];
}

// depth min/max for depth_buffer.normalize(0,255) (CImg.h:23715-23729 max_min); depths are < = 1.0f and
// finite or -inf, so the order-preserving key map of depth_key() applies (with +1.0f included).
__device__ __forceinline__ uint32_t order_key(float v) {
    uint32_t b = __float_as_uint(v);
    if (v == 0.f) b = 0u;
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float order_unkey(uint32_t k) {
    const uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return __uint_as_float(b);
}

__global__ void k_depth_minmax(const float *depth, uint32_t n, uint32_t *minmax /* [0]=min key, [1]=max key */) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = order_key(depth[i]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31u) == 0u) { atomicMin(&minmax[0], lo); atomicMax(&minmax[1], hi); }
}

// (T)((v - m)/(M - m)*(b - a) + a) with a = 0, b = 255, then the PNM writer's uchar cast
// (CImg.h:26786-26794, :52410).  m == M fills with 0.
__global__ void k_depth_to_u8(const float *depth, uint32_t n, const uint32_t *minmax, uint8_t *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using namespace exact;
    const float m = order_unkey(minmax[0]), M = order_unkey(minmax[1]);
    float v = depth[i];
    if (m == M) { out[i] = 0; return; }
    if (m != 0.f || M != 255.f) v = add(mul(div(sub(v, m), sub(M, m)), sub(255.f, 0.f)), 0.f);
    out[i] = (uint8_t)(int)v;
}

__global__ void k_count_visible(const unsigned long long *vis, uint32_t n, unsigned long long *count) {
    uint32_t c = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += vis[i] != VIS_EMPTY;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31u) == 0u && c) atomicAdd(count, (unsigned long long)c);
}

__global__ void k_count_front(Scene sc, Batch bt, uint32_t f, unsigned long long *count) {
    uint32_t c = 0;
    const float4 *rv = bt.rv + (size_t)f * sc.V;
    const bool cw = bt.frames[f].wind_clockwise != 0u;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < sc.T; t += (uint64_t)gridDim.x * blockDim.x) {
        const float a2 = signed_area_2d(rv[sc.vidx0[t]], rv[sc.vidx1[t]], rv[sc.vidx2[t]]);
        c += ((a2 > 0.f) != cw);
    }
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31u) == 0u && c) atomicAdd(count, (unsigned long long)c);
}

} // namespace rk
