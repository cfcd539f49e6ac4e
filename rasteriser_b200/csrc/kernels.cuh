// kernels.cuh -- sm_100a kernels of the frame path (vertex -> setup/cull -> rasterise with a 64-bit
// visibility buffer -> deferred resolve + shade).  Semantics follow the reference line by line
// (citations per kernel); the structure does not: the reference walks triangles one after another
// and shades every depth-passing fragment immediately (drawing.cpp:250-257, :119-146), this
// pipeline resolves visibility order-independently with atomicMin on (depth key << 32 | triangle
// index) and shades each pixel once.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cooperative_groups/scan.h>

#include "exact.cuh"

// k_setup drops the provably empty border columns / rows of a small bbox before walking it (tight_bbox.h: proof, CPU brute
// force, host emulation; on a B200: k_setup 0.154 -> 0.128 ms on 8.0 M triangles, 0.709 -> 0.592 ms on 49.9 M, identical
// frames).  -DRAST_TIGHT_TINY=0 builds the literal walk of the reference's whole bbox.
#ifndef RAST_TIGHT_TINY
#define RAST_TIGHT_TINY 1
#endif
// Timing probes (never in a product build; outputs are wrong): k_setup without its atomics / without its pixel loops
#ifndef RAST_PROBE_NO_ATOMIC
#define RAST_PROBE_NO_ATOMIC 0
#endif
#ifndef RAST_PROBE_NO_WALK
#define RAST_PROBE_NO_WALK 0
#endif
// The shade pass reads one prepared 160-byte record per (frame, triangle) -- vertices, the pixel-invariant edge differences,
// area and its refined reciprocal, depths, 1/w, camera normals, uvs, material -- written once per batch by k_prepare_tris,
// instead of gathering record -> vertices / normals / uvs and recomputing the differences for every pixel (about 40 of the
// ~330 instructions of a covered pixel and one level of dependent loads).  Hoisting rounded values does not change them, so
// the bits are the same (tests/test_emu_device_fns.py runs both flavours against the oracle; on a B200 the 1080p spin step's
// shade pass 1.44 -> 1.33 ms per 120 frames with identical frames).  The host enables it per call where the records fit
// beside the visibility buffer and the extra launch pays (draw_frames_impl).  -DRAST_SHADE_PREP=0 builds without it.
#ifndef RAST_SHADE_PREP
#define RAST_SHADE_PREP 1
#endif
#if RAST_TIGHT_TINY
#include "tight_bbox.h"
#endif

// Block-level early depth rejection in the chunk rasteriser (raster_item<.., BLOCKZ = true>), taken by k_raster_chunks for
// batches with high overdraw (the early_z condition): on a B200 the 8K overdraw frame's raster pass 5.69 -> 5.04 ms with
// identical frames; the low-overdraw instantiation is the plain loop (the variant's extra registers and tests cost the 1080p
// spin batch 4 % when it was compiled into the one loop).  -DRAST_BLOCK_Z=0 builds without it.
#ifndef RAST_BLOCK_Z
#define RAST_BLOCK_Z 1
#endif
// Device-only primitives used inside the RAST_HD functions, with host stand-ins for tests/emu_device_fns.cu.  By default a host
// "warp" is one lane at a time (a vote is the lane's own predicate, a warp maximum the lane's own value, an atomic min a plain
// min); the test may define RAST_HOST_ANY / RAST_HOST_WARP_MAX_U32 before including this file to run 32 lanes in lockstep with
// real votes and reductions.
#ifndef RAST_HOST_ANY
#define RAST_HOST_ANY(p) (p)
#endif
#ifndef RAST_HOST_WARP_MAX_U32
#define RAST_HOST_WARP_MAX_U32(v) (v)
#endif
#ifdef __CUDA_ARCH__
#define RAST_WARP_MAX_U32(v) __reduce_max_sync(0xFFFFFFFFu, (v))
#else
#define RAST_WARP_MAX_U32(v) RAST_HOST_WARP_MAX_U32(v)
#endif
#ifdef __CUDA_ARCH__
#define RAST_ANY(p) __any_sync(0xFFFFFFFFu, (p))
#define RAST_BALLOT(p) __ballot_sync(0xFFFFFFFFu, (p))
#define RAST_ATOMIC_MIN64(ptr, v) atomicMin((ptr), (v))
#define RAST_LDCG32(ptr) __ldcg(ptr)
#else
#define RAST_ANY(p) RAST_HOST_ANY(p)
#define RAST_BALLOT(p) 0u /* only the tile schedule votes across lanes; the host emulation drives the chunk schedule */
#define RAST_ATOMIC_MIN64(ptr, v) do { unsigned long long *p__ = (ptr); const unsigned long long v__ = (v); if (v__ < *p__) *p__ = v__; } while (0)
#define RAST_LDCG32(ptr) (*(ptr))
#endif

namespace rk {

constexpr unsigned long long VIS_EMPTY = ~0ull;
constexpr uint32_t INVALID_TRI = 0xFFFFFFFFu;
constexpr int CHUNK = 32;                   // a queued work item covers <= CHUNK x CHUNK pixels of a triangle's bbox
// Bbox size (pixels) up to which the setup thread rasterises a triangle itself instead of queueing it.  Measured on
// B200: 8 M-triangle mesh at 4K: 16 -> 0.64 ms/frame, 32 -> 0.42, 64 -> 0.39 (queueing a small bbox costs more than
// walking it); a 968-triangle frame at 640x480 prefers 16 (queued bboxes run in parallel).  The context picks
// clamp(T / 16384, TINY_MIN_PIXELS, TINY_MAX_PIXELS) at upload.
constexpr uint32_t TINY_MIN_PIXELS = 16;
constexpr uint32_t TINY_MAX_PIXELS = 64;
constexpr float EDGE_SLACK = 2.384185791015625e-07f; // 2^-22, see candidate()

// ---- device-side data ----------------------------------------------------------------------
struct FrameParams {          // one per frame of a batch, built on the host (hostmath.h)
    float camera[16];         // perspective * view * model   (drawing.cpp:229)
    float normal_m[16];       // transpose(inverse(modelview)) (geometry.cpp:101)
    uint32_t wind_clockwise;  // arguments.h:14
    uint32_t flat_face;       // extension (rast_args.flat == RAST_FLAT_FACE): one normal per face
    uint32_t pad[2];
    float modelview[16];      // view * model (drawing.cpp:226); only read in flat_face mode
};

struct LightDev {             // pre-combined per light: -trans_dir and intensity*colour (shading.cpp:21)
    float ntx, nty, ntz, icr, icg, icb, pad0, pad1;
};

struct MaterialDev {          // material.h:11-25
    float kd[3];
    int has_texture;          // 0 = Kd, 1 = texture (the reference), 1 | 2 = texture x Kd (extension RAST_TEXTURE_MODULATE_KD)
    int tex_w, tex_h;
    long long texel_offset;   // into Scene::texels, in texels
};

struct Scene {
    const float *pos;         // xyz [V]
    const float *nrm;         // xyz [Nn] (k_vertex: coalesced)
    const float4 *nrm4;       // the same normals padded to 16 bytes (x, y, z, 0): the shade pass of a huge mesh gathers three per pixel --
                              // one 16-byte load each instead of three scattered 4-byte ones (k_pad_normals at upload)
    const float2 *uv;         // [Nuv]
    const int *vidx0, *vidx1, *vidx2; // vertex indices, SoA [T] (coalesced in the per-triangle pass)
    const int4 *tri_rec;      // [3T]: (v0,v1,v2,n0) (n1,n2,t0,t1) (t2,material,-,-): one 48-byte record per
                              // triangle, read only by the shade pass for winning triangles.  Absent (-1)
                              // normal / uv / material indices are remapped on upload to a sentinel entry
                              // appended to each array (zero normal, uv (0,0), white untextured material),
                              // so the shade pass needs no index checks
    const MaterialDev *mats;
    const float4 *texels;     // all textures, interleaved (r, g, b, -) per texel
    uint32_t V, Nn, Nuv, M;   // Nn, Nuv, M count the appended sentinel entry
    uint64_t T;
};

struct View {
    uint32_t W, H;            // image size (arguments.h:8-9)
    uint32_t y0, y1;          // band of rows rendered by this context, [y0,y1)
    uint32_t band_pixels;     // W * (y1 - y0)
    uint32_t out_plane;       // pixels between the R, G and B planes (and between frames' depth planes) of the OUTPUT: band_pixels, or the
                              // whole frame's W * H when a band is written straight into a full-size image (rast_set_output_plane_stride)
    uint32_t out_frame_stride;// frame slots between consecutive frames of a call in the OUTPUT (1 = dense; N = every N-th slot of a sequence
                              // buffer that N ranks fill round-robin, rast_set_output_frame_stride)
};

struct Batch {
    const FrameParams *frames;
    uint32_t n_frames;
    float4 *rv;               // raster vertices (x, y, ndc z, 1/w) [n_frames][V]  (drawing.cpp:216,247)
    float4 *cn;               // camera-space normals [n_frames][Nn] (drawing.cpp:219,236) or nullptr: then the
                              // shade pass transforms the three normals of each winning triangle itself
    unsigned long long *vis;  // visibility buffer [n_frames][band_pixels]
    uint2 *queue;             // work items (triangle, cx | cy<<12 | frame<<24)
    uint32_t queue_cap;
    uint32_t tiny_max_pixels; // bboxes up to this many pixels are rasterised by the setup thread itself
    unsigned long long *counters; // [0] queue count (may exceed queue_cap), [1] queue cursor, [2] overflow flag
    unsigned long long tile_min_area; // screen-tile schedule requested for this batch (~0ull: not requested): it is taken if no bin overflowed and the
                              // queued bbox area reaches this many pixels (the host asked for bins because the PREVIOUS call showed high overdraw; a
                              // batch that turns out not to have it keeps the chunk queue, which k_setup fills as well) -- tile_schedule_taken()
    uint8_t *tile_flags;      // [n_frames][tiles_y][tiles_x]: 1 = some pass may have written a key into that 32 x 16 pixel tile (k_resolve_shade)
    uint32_t *spans;          // [n_frames][rows][2] or nullptr: covered span of every row of every frame, written by the shade pass as
                              // atomicMin of (x, W-1-x) over the row's covered pixels (0xFFFFFFFF = nothing covered in that row): only the
                              // spans travel to the caller's host buffers (k_deliver or 2-D copies), the host fills the rest itself
#if RAST_SHADE_PREP
    float4 *prep;             // prepared shading records [n_frames][T][PREP_QUADS] or nullptr (k_prepare_tris -> k_resolve_shade)
#endif
};

// Screen-tile bins of the binned raster schedule (k_setup<bins> fills them, k_raster_tiles consumes them).
struct TileBins {
    uint32_t *fill;        // [n_frames * tiles] items appended to each tile's bin so far (zeroed per batch; may exceed cap)
    uint2 *items;          // [n_frames * tiles][cap]: (triangle id, depth key of its nearest vertex), appended by k_setup
    uint32_t cap;          // slots per bin; a bin that would need more sets CNT_TILE_OVERFLOW and the batch falls back to the chunk queue
    uint32_t tiles_x, tiles_y; // tiles of one frame (band)
};

constexpr uint32_t TILE = 32; // screen tile edge of the binned schedule (= CHUNK: both use the same 4 x 2 block grid)
enum { CNT_QUEUE = 0, CNT_CURSOR = 1, CNT_QUEUE_OVERFLOW = 2, CNT_BBOX_AREA = 3, CNT_TILE_MODE = 6, CNT_TILE_OVERFLOW = 7, CNT_COUNT = 8 }; // counters[] slots

// Both raster kernels ask this once k_setup has finished: does the batch go through the screen-tile bins (k_raster_tiles) or the chunk queue?
__device__ __forceinline__ bool tile_schedule_taken(const Batch &bt) {
    return bt.tile_min_area != ~0ull && bt.counters[CNT_TILE_OVERFLOW] == 0ull && bt.counters[CNT_BBOX_AREA] >= bt.tile_min_area;
}

// ---- visibility key ------------------------------------------------------------------------
// Order-preserving map of a finite float below 1.0 to u32.  -0 is canonicalised so that it ties
// with +0 (the reference's strict '<' treats them as equal, drawing.cpp:119).
RAST_HD uint32_t depth_key(float z) {
    uint32_t b = exact::f2u(z);
    if (z == 0.f) b = 0u;
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// ---- triangle setup ------------------------------------------------------------------------
struct TriSetup {
    float x0, y0, x1, y1, x2, y2;
    float z0, z1, z2;
    float d12x, d12y, d20x, d20y, d01x, d01y; // the pixel-invariant differences of edge() (drawing.cpp:38), times sign(area)
    float area;                               // |edge(v2; v0, v1)| (drawing.cpp:46)
    float rcp1;                               // exact::div_reciprocal(area), for the shared-divisor quotients
    bool div_ok;                              // exact::div_in_range(area)
    bool literal;                             // area is 0, inf or NaN: no sign shortcut
};

// The six differences and the area are multiplied by sign(area).  Negation is exact and every IEEE
// operation is sign-symmetric, so each edge value becomes exactly sign(area) * e_k and each quotient
// e_k / area keeps its bits; what changes is that "inside" now simply reads e_k >= 0.
RAST_HD void tri_setup(TriSetup &s, const float4 &v0, const float4 &v1, const float4 &v2) {
    s.x0 = v0.x; s.y0 = v0.y; s.z0 = v0.z;
    s.x1 = v1.x; s.y1 = v1.y; s.z1 = v1.z;
    s.x2 = v2.x; s.y2 = v2.y; s.z2 = v2.z;
    s.d12x = exact::sub(v2.x, v1.x); s.d12y = exact::sub(v2.y, v1.y);
    s.d20x = exact::sub(v0.x, v2.x); s.d20y = exact::sub(v0.y, v2.y);
    s.d01x = exact::sub(v1.x, v0.x); s.d01y = exact::sub(v1.y, v0.y);
    s.area = exact::sub(exact::mul(s.d01x, exact::sub(v2.y, v0.y)), exact::mul(s.d01y, exact::sub(v2.x, v0.x)));
    s.literal = !(fabsf(s.area) > 0.f && fabsf(s.area) < exact::i2f(0x7f800000));
    if (!s.literal && s.area < 0.f) {
        s.d12x = -s.d12x; s.d12y = -s.d12y; s.d20x = -s.d20x; s.d20y = -s.d20y; s.d01x = -s.d01x; s.d01y = -s.d01y;
        s.area = -s.area;
    }
    s.div_ok = exact::div_in_range(s.area);
    s.rcp1 = exact::div_reciprocal(s.area);
}

// signed_area_2d (geometry.cpp:76-83), left to right
RAST_HD float signed_area_2d(const float4 &v0, const float4 &v1, const float4 &v2) {
    float a = exact::sub(exact::mul(v0.x, v1.y), exact::mul(v1.x, v0.y));
    a = exact::add(a, exact::mul(v1.x, v2.y));
    a = exact::sub(a, exact::mul(v2.x, v1.y));
    a = exact::add(a, exact::mul(v2.x, v0.y));
    a = exact::sub(a, exact::mul(v0.x, v2.y));
    return exact::mul(-0.5f, a);
}

struct BBox { uint32_t x0, y0, x1, y1; bool empty; };

// bounding_box (drawing.cpp:77-93) intersected with the band [vy0, vy1)
RAST_HD BBox bounding_box(const float4 &v0, const float4 &v1, const float4 &v2, const View &vw) {
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
RAST_HD void edges(const TriSetup &s, float px, float py, float &e0, float &e1, float &e2) {
    using namespace exact;
    e0 = sub(mul(s.d12x, sub(py, s.y1)), mul(s.d12y, sub(px, s.x1)));
    e1 = sub(mul(s.d20x, sub(py, s.y2)), mul(s.d20y, sub(px, s.x2)));
    e2 = sub(mul(s.d01x, sub(py, s.y0)), mul(s.d01y, sub(px, s.x0)));
}

// Cheap superset of the inside test.  The reference tests e_k/area >= 0 (drawing.cpp:46-48,111).
// For finite non-zero area the quotient is >= 0 (counting -0) iff e_k has area's sign, is zero, or
// the quotient underflows to -0; underflow needs |e_k| <= 2^-150 * |area| < 2^-22.  With the setup's
// sign folding that is "e_k >= -2^-22 for all k": it never rejects a pixel the exact test accepts
// (fminf drops a NaN operand, which only widens the superset); survivors take the literal divisions,
// which the depth needs anyway.
RAST_HD bool candidate(const TriSetup &s, float e0, float e1, float e2) {
    return s.literal || fminf(fminf(e0, e1), e2) >= -EDGE_SLACK;
}

// barycentric + inside + depth (drawing.cpp:41-49,111,115-119).  True iff the fragment is inside
// and nearer than the cleared depth 1.0f (a fragment at z >= 1 or NaN can never pass the strict '<').
RAST_HD bool fragment(const TriSetup &s, float e0, float e1, float e2, float &b0, float &b1, float &b2, float &z) {
    using namespace exact;
    div3(e0, e1, e2, s.area, s.rcp1, s.div_ok, b0, b1, b2);
    if (!(b0 >= 0.f && b1 >= 0.f && b2 >= 0.f)) return false;
    z = add(add(mul(s.z0, b0), mul(s.z1, b1)), mul(s.z2, b2));
    return z < 1.0f;
}

RAST_HD void test_and_commit(const TriSetup &s, uint32_t x, uint32_t y, uint32_t tri, unsigned long long *vis_row0, const View &vw) {
    float e0, e1, e2, b0, b1, b2, z;
    edges(s, (float)x, (float)y, e0, e1, e2);
    if (!candidate(s, e0, e1, e2)) return;
    if (!fragment(s, e0, e1, e2, b0, b1, b2, z)) return;
    const unsigned long long key = ((unsigned long long)depth_key(z) << 32) | tri;
#if RAST_PROBE_NO_ATOMIC
    if (key != 0ull) return; // timing probe only (never true for a real key): everything but the atomic
#endif
    RAST_ATOMIC_MIN64(vis_row0 + (size_t)(y - vw.y0) * vw.W + x, key);
}

// ---- K0: clear ------------------------------------------------------------------------------
// renderer.cpp:85-86 / :107-108 (frame = 0, depth = 1.0f) become "no triangle" in the visibility buffer.
// Only needed for slots that are not known to be empty: the shade pass hands every key it consumes
// back as VIS_EMPTY, so in steady state the buffer is already clear when the next batch starts.
// (The 16-byte stores need a 16-byte aligned start; an odd slot of a band with an odd pixel count -- the one kept slot of
// rast_set_keep_visibility, 641 x 483 -- starts 8 bytes off: its first key is written on its own.)
__global__ void k_clear(unsigned long long *vis, size_t n) {
    const size_t head = (reinterpret_cast<uintptr_t>(vis) & 8u) ? 1u : 0u;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (head && t == 0 && n) vis[0] = VIS_EMPTY;
    const size_t i = head + t * 2;
    if (i + 1 < n) {
        *reinterpret_cast<ulonglong2 *>(vis + i) = make_ulonglong2(VIS_EMPTY, VIS_EMPTY);
    } else if (i < n) {
        vis[i] = VIS_EMPTY;
    }
}

// ---- K1: vertex stage -----------------------------------------------------------------------
// transform_point + z_divide + ndc_to_raster (geometry.cpp:44-74, drawing.cpp:241-247) fused: one
// thread per vertex, 12 B in, one float4 out.
RAST_HD float4 raster_vertex(const float *cam, float x, float y, float z, uint32_t W, uint32_t H) {
    using namespace exact;
    const float4 clip = mat_vec(cam, x, y, z, 1.f);
    const float nx = div(clip.x, clip.w), ny = div(clip.y, clip.w), nz = div(clip.z, clip.w), nw = div(1.f, clip.w);
    float4 r;
    r.x = mul(mul(0.5f, add(nx, 1.0f)), (float)(int)W);   // remap_ndc(x, width)
    r.y = mul(mul(0.5f, add(-ny, 1.0f)), (float)(int)H);  // remap_ndc(-y, height)
    r.z = nz;
    r.w = nw;
    return r;
}

__global__ void __launch_bounds__(256) k_vertex(Scene sc, View vw, Batch bt) {
    __shared__ float cam[16], nm[16];
    const uint32_t f = blockIdx.y;
    if (blockIdx.x == 0 && f == 0 && threadIdx.x >= 32 && threadIdx.x < 40) bt.counters[threadIdx.x - 32] = 0ull; // work queue / bin counters reset
    if (threadIdx.x < 16) cam[threadIdx.x] = bt.frames[f].camera[threadIdx.x];
    else if (threadIdx.x < 32) nm[threadIdx.x - 16] = bt.frames[f].normal_m[threadIdx.x - 16];
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    using namespace exact;
    if (i < sc.V) {
        const float x = sc.pos[3 * (size_t)i], y = sc.pos[3 * (size_t)i + 1], z = sc.pos[3 * (size_t)i + 2];
        bt.rv[(size_t)f * sc.V + i] = raster_vertex(cam, x, y, z, vw.W, vw.H);
    } else if (bt.cn != nullptr && i - sc.V < sc.Nn) {
        // transform_normals (geometry.cpp:97-108): transpose(inverse(modelview)) * (n, 0), xyz kept, not normalised
        const uint32_t j = i - sc.V;
        const float4 t = mat_vec(nm, sc.nrm[3 * (size_t)j], sc.nrm[3 * (size_t)j + 1], sc.nrm[3 * (size_t)j + 2], 0.f);
        bt.cn[(size_t)f * sc.Nn + j] = make_float4(t.x, t.y, t.z, 0.f);
    }
}

// ---- tile flags ------------------------------------------------------------------------------
// Tile flags: one byte per SHADE_TILE_W x SHADE_TILE_H (32 x 16) pixel tile of each frame of the batch, set by every pass
// that may write a key into the tile (k_setup for the triangles it rasterises itself, the raster kernels for every work item
// they stage -- conservatively: the item's rectangle, not its coverage).  A tile whose flag is 0 holds VIS_EMPTY keys only,
// so the shade pass neither reads nor resets them: it writes the cleared frame (0, 0, 0 / 1.0f: renderer.cpp:85-86) with a
// handful of 16-byte stores.  On a Suzanne frame two thirds of the tiles are like that; before, the background cost 35 % of
// the pass's instructions (40 per group of 32 pixels, ncu source counters) and all of its key reads.
constexpr uint32_t SHADE_TILE_W = 32, SHADE_TILE_H = 16, SHADE_WARPS = 4;

RAST_HD uint32_t flag_tiles_x(const View &vw) { return (vw.W + SHADE_TILE_W - 1u) / SHADE_TILE_W; }
RAST_HD uint32_t flag_tiles_y(const View &vw) { return (vw.y1 - vw.y0 + SHADE_TILE_H - 1u) / SHADE_TILE_H; }

// Sparse device-to-host delivery.  The shade pass records the covered span of every row (Batch::spans).  Where the caller's buffers are
// mapped into the device's address space k_deliver stores exactly those spans into them; elsewhere the copy engine moves one rectangle
// per horizontal STRIP of tile rows (the union of the strip's spans) -- a silhouette is much narrower near its top and bottom than its
// bounding box (Suzanne at 1080p: the box is 50 % of the frame, four strip boxes 39 %, the row spans 30 %, the covered pixels 27 %), and
// what is in no box does not cross PCIe.  Tile row ty belongs to strip ty * BBOX_STRIPS / tiles_y.
constexpr uint32_t BBOX_STRIPS = 4;
RAST_HD uint32_t bbox_strip_of_tile_row(uint32_t ty, uint32_t tiles_y) { return min(BBOX_STRIPS - 1u, ty * BBOX_STRIPS / tiles_y); }

// mark the tiles of frame f that the pixel rectangle [x0,x1] x [y0,y1] (image rows, inside the band) overlaps
__device__ __forceinline__ void mark_tiles(uint8_t *__restrict__ flags, uint32_t f, const View &vw, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) {
    const uint32_t tx_n = flag_tiles_x(vw);
    uint8_t *fl = flags + (size_t)f * tx_n * flag_tiles_y(vw);
    const uint32_t tx0 = x0 / SHADE_TILE_W, tx1 = x1 / SHADE_TILE_W, ty0 = (y0 - vw.y0) / SHADE_TILE_H, ty1 = (y1 - vw.y0) / SHADE_TILE_H;
    for (uint32_t ty = ty0; ty <= ty1; ++ty)
        for (uint32_t tx = tx0; tx <= tx1; ++tx) fl[ty * tx_n + tx] = 1;
}


// ---- K2: triangle setup, cull, classification -----------------------------------------------
// draw_triangle up to the pixel loops (drawing.cpp:165-188).  Tiny bboxes are rasterised here;
// larger ones are cut into CHUNK x CHUNK work items for k_raster_chunks.
#ifndef RAST_SETUP_TRIS
#define RAST_SETUP_TRIS 1
#endif
constexpr int SETUP_TRIS = RAST_SETUP_TRIS; // triangles per thread: all index and vertex loads of a thread are in flight together
                                            // (k_setup on 8 M / 50 M triangles, out-of-line body: 1 -> 0.187 / 0.898 ms, 2 -> 0.188 / 0.847,
                                            //  4 -> 0.225 / 1.011, 8 -> 0.323 / 1.422; inlined body: 1 -> 0.154 / 0.706, 2 -> 0.175 / 0.817.  Once the queue
                                            //  reservation is one atomic per warp, more triangles per thread only cost registers.)
// With the tight bbox walk the pixel loops are short and the kernel is bound by the latency of its two dependent load stages (ncu: 58 % of
// the stall samples): two triangles per thread then pay on big meshes (8.0 M triangles: 0.128 -> 0.114 ms; 3 -> 0.120, 4 -> 0.149), while a
// 968-triangle frame loses a few microseconds -- the host picks k_setup<.., 2> from SETUP_TRIS2_MIN_TRIANGLES triangles on.
constexpr uint64_t SETUP_TRIS2_MIN_TRIANGLES = 1ull << 20;

#ifndef RAST_SETUP_INLINE
#define RAST_SETUP_INLINE __forceinline__
#endif
template <bool BINS>
__device__ RAST_SETUP_INLINE void setup_triangle(uint32_t t, uint32_t f, float4 v0, float4 v1, float4 v2, bool cw, const View &vw, const Batch &bt, const TileBins &tb) {
    const float a2 = signed_area_2d(v0, v1, v2);
    if (!((a2 > 0.f) != cw)) return; // back face (drawing.cpp:178-180)

    const BBox bb = bounding_box(v0, v1, v2, vw);
    if (bb.empty) return;
    const uint32_t w = bb.x1 - bb.x0 + 1u, h = bb.y1 - bb.y0 + 1u;
    unsigned long long *vis = bt.vis + (size_t)f * vw.band_pixels;

    bool inline_raster = (uint64_t)w * h <= bt.tiny_max_pixels;
    if (!inline_raster) {
        const uint32_t ncx = (w + CHUNK - 1) / CHUNK, ncy = (h + CHUNK - 1) / CHUNK;
        const uint64_t n = (uint64_t)ncx * ncy;
        // Reserve n consecutive slots.  One atomicAdd per warp, not per triangle: the lanes that reached this
        // point sum their requests (exclusive scan over the coalesced group) and the last lane adds the total --
        // per-lane atomics on this single counter serialised in L2 and were 29 % of the kernel (ncu, 8 M triangles).
        // The 64-bit count keeps growing past the capacity, so it cannot wrap.
        namespace cg = cooperative_groups;
        const cg::coalesced_group grp = cg::coalesced_threads();
        const unsigned long long want = n <= bt.queue_cap ? n : 0ull;
        const unsigned long long before = cg::exclusive_scan(grp, want);
        unsigned long long base = 0;
        if (grp.thread_rank() == grp.size() - 1) base = atomicAdd(&bt.counters[0 /*CNT_QUEUE*/], before + want);
        base = grp.shfl(base, grp.size() - 1);
        // queued bbox area (also one atomic per warp): the raster pass derives its overdraw estimate from it
        const unsigned long long area_sum = cg::reduce(grp, (unsigned long long)w * h, cg::plus<unsigned long long>());
        if (grp.thread_rank() == 0) atomicAdd(&bt.counters[3 /*CNT_BBOX_AREA*/], area_sum);
        const uint64_t first = (n <= bt.queue_cap) ? base + before : (uint64_t)bt.queue_cap;
        if (BINS) { // binned schedule requested: the triangle goes straight into the bin of every tile its bbox touches (one pass: a slot is an atomicAdd away)
            const uint32_t zkey = depth_key(fminf(fminf(v0.z, v1.z), v2.z)); // nearest vertex: the tile's CTA processes its bin near to far (order only)
            const uint32_t tx0 = bb.x0 / TILE, tx1 = bb.x1 / TILE, ty0 = (bb.y0 - vw.y0) / TILE, ty1 = (bb.y1 - vw.y0) / TILE;
            const size_t g0 = (size_t)f * tb.tiles_x * tb.tiles_y;
            for (uint32_t ty = ty0; ty <= ty1; ++ty)
                for (uint32_t tx = tx0; tx <= tx1; ++tx) {
                    const size_t g = g0 + (size_t)ty * tb.tiles_x + tx;
                    const uint32_t pos = atomicAdd(&tb.fill[g], 1u);
                    if (pos < tb.cap) tb.items[g * tb.cap + pos] = make_uint2(t, zkey);
                    else bt.counters[CNT_TILE_OVERFLOW] = 1ull;
                }
        }
        if (first + n > bt.queue_cap) {
            // queue full: void any slots reserved below the capacity and walk the whole bbox in this
            // thread (correct, slow); the host sees the flag and grows the queue for later frames
            for (uint64_t k = first; k < bt.queue_cap; ++k) bt.queue[k] = make_uint2(INVALID_TRI, 0u);
            bt.counters[2 /*CNT_QUEUE_OVERFLOW*/] = 1ull;
            inline_raster = true;
        } else {
            uint64_t k = first;
            for (uint32_t cy = 0; cy < ncy; ++cy)
                for (uint32_t cx = 0; cx < ncx; ++cx) bt.queue[k++] = make_uint2(t, cx | (cy << 12) | (f << 24));
        }
    }
    if (inline_raster) {
        TriSetup s;
        tri_setup(s, v0, v1, v2);
        // (walking the bbox as 2x2 quads with shared differences in one flat loop measured slower: 0.173 / 0.844 ms against
        //  0.156 / 0.710 ms on 8 M / 50 M triangles -- sub-pixel triangles have 2-3 pixel wide bboxes, quads test 28 % more pixels)
        uint32_t wx0 = bb.x0, wy0 = bb.y0, wx1 = bb.x1, wy1 = bb.y1;
#if RAST_TIGHT_TINY
        // border columns / rows of the bbox that lie outside the triangle's extent by more than the rounding error of the
        // reference's own test can hold no candidate pixel (tight_bbox.h: proof + CPU brute-force check); a sub-pixel
        // triangle between sample points is dropped here without a single pixel test
        if (!rast_tight_bbox(v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, s.literal ? 0.f : s.area, &wx0, &wy0, &wx1, &wy1)) return;
#endif
#if RAST_PROBE_NO_WALK
        if (wx0 != 0xFFFFFFFFu) return; // timing probe only: the per-triangle cost without any pixel test
#endif
        mark_tiles(bt.tile_flags, f, vw, wx0, wy0, wx1, wy1); // keys may appear in these tiles (k_resolve_shade skips the others)
        for (uint32_t y = wy0; y <= wy1; ++y)
            for (uint32_t x = wx0; x <= wx1; ++x) test_and_commit(s, x, y, t, vis, vw);
    }
}

#ifndef RAST_SETUP_MIN_BLOCKS
#define RAST_SETUP_MIN_BLOCKS 0 // variant: resident CTAs per SM asked of the register allocator (0 = the compiler's choice, 48 registers / 5 CTAs)
#endif
template <bool BINS, int TRIS = SETUP_TRIS>
#if RAST_SETUP_MIN_BLOCKS > 0
__global__ void __launch_bounds__(256, RAST_SETUP_MIN_BLOCKS) k_setup(
#else
__global__ void __launch_bounds__(256) k_setup(
#endif
const __grid_constant__ Scene sc, const __grid_constant__ View vw, const __grid_constant__ Batch bt,
                                               const __grid_constant__ TileBins tb) {
    const uint32_t f = blockIdx.y;
    const uint64_t t0 = (uint64_t)blockIdx.x * (256 * TRIS) + threadIdx.x;
    // The frame's vertices behind an opaque base, UNSIGNED 32-bit indices (validated at upload) and read-only loads (k_vertex wrote them in
    // an earlier launch): a gather is IMAD.WIDE + LDG instead of a sign extension, a 64-bit multiply-add of the frame offset and a LEA pair
    unsigned long long rv_base = (unsigned long long)(bt.rv + (size_t)f * sc.V);
    asm volatile("" : "+l"(rv_base));
    const float4 *rv = reinterpret_cast<const float4 *>(rv_base);
    // phase 1: the (coalesced) index loads of all this thread's triangles, then all the vertex gathers
    uint32_t i0[TRIS], i1[TRIS], i2[TRIS];
#pragma unroll
    for (int k = 0; k < TRIS; ++k) {
        const uint64_t t = t0 + (uint64_t)k * 256;
        const bool in = t < sc.T;
        i0[k] = in ? (uint32_t)sc.vidx0[t] : 0u; i1[k] = in ? (uint32_t)sc.vidx1[t] : 0u; i2[k] = in ? (uint32_t)sc.vidx2[t] : 0u;
    }
    float4 v0[TRIS], v1[TRIS], v2[TRIS];
#pragma unroll
    for (int k = 0; k < TRIS; ++k) { v0[k] = exact::ldg(rv + i0[k]); v1[k] = exact::ldg(rv + i1[k]); v2[k] = exact::ldg(rv + i2[k]); }
    const bool cw = bt.frames[f].wind_clockwise != 0u;
    // phase 2
#pragma unroll
    for (int k = 0; k < TRIS; ++k) {
        const uint64_t t = t0 + (uint64_t)k * 256;
        if (t < sc.T) setup_triangle<BINS>((uint32_t)t, f, v0[k], v1[k], v2[k], cw, vw, bt, tb);
    }
}

// The same pass as a software pipeline, for meshes far larger than the L2.  k_setup above is bound by the latency of its two dependent
// load stages (ncu, 8.0 M triangles: 35 % of the stall samples sit on the index loads and on the vertex gathers, issue slots 69 % used):
// every warp loads, waits, gathers, waits, computes.  Here a persistent CTA walks chunks of 256 consecutive triangles (chunk c, c + grid,
// ...) and every thread keeps two loads ahead of its arithmetic: while triangle i is set up, the vertices of triangle i + 1 and the
// indices of triangle i + 2 are in flight.  Same per-triangle code, same results.  Measured on a B200 against k_setup<.., 2>: 49.9 M
// triangles (1.0 GB of indices and vertices, DRAM latency) 0.615 -> 0.563 ms; 8.0 M triangles (160 MB, mostly L2 hits after k_vertex)
// 0.133 -> 0.169 ms (one wave per SM) / 0.222 ms -- the host takes it from SETUP_PIPE_MIN_TRIANGLES triangles on.
constexpr uint64_t SETUP_PIPE_MIN_TRIANGLES = 1ull << 25;
template <bool BINS>
__global__ void __launch_bounds__(256) k_setup_pipe(const __grid_constant__ Scene sc, const __grid_constant__ View vw, const __grid_constant__ Batch bt,
                                                    const __grid_constant__ TileBins tb) {
    const uint32_t f = blockIdx.y;
    unsigned long long rv_base = (unsigned long long)(bt.rv + (size_t)f * sc.V); // (opaque base + unsigned indices + read-only loads, as in k_setup)
    asm volatile("" : "+l"(rv_base));
    const float4 *rv = reinterpret_cast<const float4 *>(rv_base);
    const bool cw = bt.frames[f].wind_clockwise != 0u;
    const uint64_t stride = (uint64_t)gridDim.x * 256u;
    uint64_t t = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    // prologue: indices of the first two triangles, vertices of the first
    uint32_t a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
    if (t < sc.T) { a0 = (uint32_t)sc.vidx0[t]; a1 = (uint32_t)sc.vidx1[t]; a2 = (uint32_t)sc.vidx2[t]; }
    if (t + stride < sc.T) { b0 = (uint32_t)sc.vidx0[t + stride]; b1 = (uint32_t)sc.vidx1[t + stride]; b2 = (uint32_t)sc.vidx2[t + stride]; }
    float4 v0 = exact::ldg(rv + a0), v1 = exact::ldg(rv + a1), v2 = exact::ldg(rv + a2);
#pragma unroll 1
    for (; t < sc.T; t += stride) {
        // vertices of the next triangle (its indices arrived during the previous iteration), indices of the one after
        const float4 n0 = exact::ldg(rv + b0), n1 = exact::ldg(rv + b1), n2 = exact::ldg(rv + b2);
        uint32_t c0 = 0, c1 = 0, c2 = 0;
        const uint64_t t2 = t + 2 * stride;
        if (t2 < sc.T) { c0 = (uint32_t)sc.vidx0[t2]; c1 = (uint32_t)sc.vidx1[t2]; c2 = (uint32_t)sc.vidx2[t2]; }
        setup_triangle<BINS>((uint32_t)t, f, v0, v1, v2, cw, vw, bt, tb);
        v0 = n0; v1 = n1; v2 = n2;
        b0 = c0; b1 = c1; b2 = c2;
    }
}

// ---- K3: rasterisers --------------------------------------------------------------------------
// update_pixel's coverage + depth part (drawing.cpp:108-121) for the queued triangles.  Two schedules share
// one inner loop:
//   * k_raster_chunks -- work items are <=32x32-pixel chunks of a triangle's bbox, taken from a flat queue in
//     any order; fragments go to the visibility buffer with global atomicMin.  Best when few fragments
//     compete per pixel.
//   * k_raster_tiles  -- screen-tile binning: the items of one aligned 32x32 screen tile are processed by one
//     CTA that keeps the tile's 1024 keys in shared memory (atomicMin + early depth rejection against shared
//     memory) and merges them into the visibility buffer once.  Chosen by the host when the previous call
//     showed high overdraw: the global early-z reads of the chunk schedule then miss L2 and dominate.
// In both, a warp stages up to 32 items at a time: each lane fetches ONE item (indices -> vertices: the
// dependent loads overlap across the lanes), computes its triangle setup, the conservative block mask and
// the early-z constants, and writes them to shared memory; then the whole warp walks the staged items.
// Lanes own 2x2 pixel quads of a 16x8 block, so the per-pixel differences (p - a) and half of the products
// of edge() are shared inside the quad; a ballot skips blocks no lane may cover.
constexpr int RASTER_WARPS = 8;
constexpr int STAGE_FIELDS = RAST_BLOCK_Z ? 31 : 27;
constexpr unsigned long long EARLY_Z_OVERDRAW = 6;

// Conservative rejection of one 16x8 block (pixel extent [xa,xb] x [ya,yb]) against one sign-folded edge.
// E(p) = dx*(py-yk) - dy*(px-xk) is affine, so over the block it peaks at a corner pixel; the fp32
// evaluation e(p) differs from E(p) by at most 3.01 ulp-units of (|dx|*|py-yk| + |dy|*|px-xk|) (two rounded
// differences, two rounded products, one rounded subtraction) plus 2^-148 for underflow.  Hence
//   e(p) <= max_corners e(c) + 2*err   for every pixel p of the block,
// and if that bound is below the candidate threshold -2^-22 no pixel of the block can be a candidate.
// 8 ulp-units (2^-21) are used for 2*err = 6.02; NaN / inf make the comparison false (no rejection).
RAST_HD bool block_outside_edge(float dx, float dy, float xk, float yk, float xa, float xb, float ya, float yb) {
    using namespace exact;
    const float ua = sub(ya, yk), ub = sub(yb, yk), va = sub(xa, xk), vb = sub(xb, xk);
    const float pa = mul(dx, ua), pb = mul(dx, ub), qa = mul(dy, va), qb = mul(dy, vb);
    const float emax = fmaxf(fmaxf(sub(pa, qa), sub(pa, qb)), fmaxf(sub(pb, qa), sub(pb, qb)));
    const float spread = fabsf(dx) * fmaxf(fabsf(ua), fabsf(ub)) + fabsf(dy) * fmaxf(fabsf(va), fabsf(vb));
    const float err = fmaf(spread, 4.76837158203125e-07f /* 2^-21 */, 7.17e-43f /* 2^-140 */);
    return emax + err < -EDGE_SLACK;
}

struct StagedTris {
    uint32_t w[STAGE_FIELDS][32]; // [field][slot]: conflict-free lane-per-slot writes, broadcast reads
};

// Stage one item (this lane's) : triangle `tri` of frame `f`, to be rasterised inside the pixel rectangle
// [rx0,rx1] x [ry0,ry1] (already intersected with its bbox) on the 4 x 2 block grid anchored at (ox, oy).
RAST_HD void stage_item(StagedTris &stg, uint32_t lane, uint32_t tri, uint32_t f, const float4 &v0, const float4 &v1, const float4 &v2,
                                           uint32_t rx0, uint32_t ry0, uint32_t rx1, uint32_t ry1, uint32_t ox, uint32_t oy) {
    TriSetup s = {};
    s.literal = true;
    if (tri != INVALID_TRI) tri_setup(s, v0, v1, v2);
    const float fl[16] = {s.x0, s.y0, s.x1, s.y1, s.x2, s.y2, s.z0, s.z1, s.z2, s.d12x, s.d12y, s.d20x, s.d20y, s.d01x, s.d01y, s.area};
#pragma unroll
    for (int k = 0; k < 16; ++k) stg.w[k][lane] = exact::f2u(fl[k]);
    stg.w[16][lane] = s.literal ? 1u : 0u;
    stg.w[17][lane] = rx0 | (ry0 << 16);
    stg.w[18][lane] = rx1 | (ry1 << 16);
    stg.w[19][lane] = tri;
    stg.w[20][lane] = f;
    // early depth rejection (see raster_item): 1/area and an error margin, both only used to SKIP work
    const float zmax = fmaxf(fmaxf(fabsf(s.z0), fabsf(s.z1)), fabsf(s.z2));
    const float rcp = 1.0f / s.area;
    const bool usable = !s.literal && rcp > 0.f && rcp < exact::i2f(0x7f800000) && zmax < exact::i2f(0x7f800000);
    stg.w[21][lane] = exact::f2u(usable ? rcp : 0.f);
    stg.w[22][lane] = exact::f2u(usable ? fmaf(zmax, 1.9073486328125e-06f /* 2^-19 */, 1e-37f) : exact::i2f(0x7f800000));
#if RAST_BLOCK_Z
    // Depth plane of the item for the block-level rejection in raster_item: z as an affine function of the pixel, anchored at the
    // rectangle's first pixel, and a margin M that covers every rounding between this plane and the depth the exact path would
    // compute for an ACCEPTED pixel p of the rectangle:  z(p) >= Zo + gx (p.x - rx0) + gy (p.y - ry0) - M.
    // Derivation (u = 2^-24, A = folded area, S = wt DY + ht DX as in tight_bbox.h, Q = (S + wt ht + 32 (wt + ht)) / A):
    //   exact z(p) against the real-arithmetic plane: numerators off by <= errE ~ 4uS, area by errA ~ 8u wt ht, three
    //   divisions, five rounded operations, accepted pixels have 0 <= b_k <= 1 + 3 errE / A       -> zmax u (108.3 Q + 16)
    //   Zo = z_est at the anchor, which may lie far outside the triangle (|e_k| <= S)                -> zmax u (29.8 Q + 26.5 Q^2)
    //   the two gradients (fma chains of z_k d_k, times 1/A), over at most 31 pixels each            -> zmax u (39.6 Q + 53 Q^2)
    //   evaluating the plane in fp32 at a corner                                                      -> zmax u 20 Q
    // M = zmax u (256 Q + 128 Q^2 + 32) + 1e-30 bounds the sum with room to spare; ill-conditioned items (Q > 1000, unusable
    // reciprocal, non-finite depths) get M = inf, i.e. are never rejected by the block test.  Only used to SKIP work.
    {
        const float minx = fminf(fminf(s.x0, s.x1), s.x2), maxx = fmaxf(fmaxf(s.x0, s.x1), s.x2);
        const float miny = fminf(fminf(s.y0, s.y1), s.y2), maxy = fmaxf(fmaxf(s.y0, s.y1), s.y2);
        const float wt = maxx - minx, ht = maxy - miny;
        const float fx0 = (float)rx0, fx1 = (float)rx1, fy0 = (float)ry0, fy1 = (float)ry1;
        const float DX = fmaxf(fx1 - minx, maxx - fx0), DY = fmaxf(fy1 - miny, maxy - fy0);
        const float Q = (wt * DY + ht * DX + wt * ht + 32.f * (wt + ht)) * rcp;
        float e0, e1, e2;
        edges(s, fx0, fy0, e0, e1, e2);
        const float Zo = fmaf(s.z2, e2, fmaf(s.z1, e1, s.z0 * e0)) * rcp;
        const float gx = -fmaf(s.z2, s.d01y, fmaf(s.z1, s.d20y, s.z0 * s.d12y)) * rcp; // d e_k / d px = -d_ky
        const float gy = fmaf(s.z2, s.d01x, fmaf(s.z1, s.d20x, s.z0 * s.d12x)) * rcp;  // d e_k / d py = +d_kx
        const bool ok = usable && Q <= 1000.f && Zo == Zo && fabsf(Zo) < exact::i2f(0x7f800000) && fabsf(gx) < exact::i2f(0x7f800000) && fabsf(gy) < exact::i2f(0x7f800000);
        const float M = ok ? fmaf(zmax * 5.9604644775390625e-08f /* u */, fmaf(Q, fmaf(Q, 128.f, 256.f), 32.f), 1e-30f) : exact::i2f(0x7f800000);
        stg.w[27][lane] = exact::f2u(ok ? Zo : 0.f);
        stg.w[28][lane] = exact::f2u(ok ? gx : 0.f);
        stg.w[29][lane] = exact::f2u(ok ? gy : 0.f);
        stg.w[30][lane] = exact::f2u(M);
    }
#endif
    // which of the 4 x 2 blocks can contain a candidate pixel at all (bit = strip * 2 + column)
    uint32_t live = 0u;
    if (tri != INVALID_TRI) {
#pragma unroll
        for (uint32_t b = 0; b < 8u; ++b) {
            const uint32_t bx0 = max(ox + (b & 1u) * 16u, rx0), by0 = max(oy + (b >> 1) * 8u, ry0);
            const uint32_t bx1 = min(ox + (b & 1u) * 16u + 15u, rx1), by1 = min(oy + (b >> 1) * 8u + 7u, ry1);
            if (bx0 > bx1 || by0 > by1) continue;
            const float xa = (float)bx0, xb = (float)bx1, ya = (float)by0, yb = (float)by1;
            if (!s.literal && (block_outside_edge(s.d12x, s.d12y, s.x1, s.y1, xa, xb, ya, yb) || block_outside_edge(s.d20x, s.d20y, s.x2, s.y2, xa, xb, ya, yb) ||
                               block_outside_edge(s.d01x, s.d01y, s.x0, s.y0, xa, xb, ya, yb)))
                continue;
            live |= 1u << b;
        }
    }
    stg.w[23][lane] = live;
    stg.w[24][lane] = ox | (oy << 16);
    stg.w[25][lane] = exact::f2u(s.rcp1);
    stg.w[26][lane] = s.div_ok ? 1u : 0u;
}

// Block-level depth rejection of the tile schedule: does the item's depth plane (anchored at its rectangle's first pixel, stage_item), less
// its proven margin, lie behind the farthest depth `kmax` (a depth key) stored in 16 x 8 block b of the tile?  Then every fragment of the
// item in that block would lose its atomicMin.  kmax = 0xFFFFFFFF: some pixel of the block is still empty, nothing is rejected.
RAST_HD bool block_behind(uint32_t b, uint32_t kmax, uint32_t rx0, uint32_t ry0, uint32_t rx1, uint32_t ry1, uint32_t ox, uint32_t oy,
                          float bz_o, float bz_gx, float bz_gy, float bz_m) {
    if (kmax == 0xFFFFFFFFu) return false;
    const uint32_t bs = b >> 1, bc = b & 1u;
    const uint32_t bxa = max(ox + bc * 16u, rx0), bxb = min(ox + bc * 16u + 15u, rx1);
    const uint32_t bya = max(oy + bs * 8u, ry0), byb = min(oy + bs * 8u + 7u, ry1);
    const float dxa = (float)(bxa - rx0), dxb = (float)(bxb - rx0), dya = (float)(bya - ry0), dyb = (float)(byb - ry0);
#ifdef RAST_BLOCK_Z_TEST_BIAS // (tests/test_emu_device_fns.py builds with a positive bias to show that a wrong bound is caught)
    const float lb = bz_o + fminf(bz_gx * dxa, bz_gx * dxb) + fminf(bz_gy * dya, bz_gy * dyb) - bz_m + RAST_BLOCK_Z_TEST_BIAS;
#else
    const float lb = bz_o + fminf(bz_gx * dxa, bz_gx * dxb) + fminf(bz_gy * dya, bz_gy * dyb) - bz_m;
#endif
    const uint32_t far_bits = (kmax & 0x80000000u) ? (kmax ^ 0x80000000u) : ~kmax; // inverse of depth_key
    return lb > exact::u2f(far_bits);
}

// Rasterise staged item `it` with the whole warp.  TILE_MODE = false: keys go to the visibility buffer `vis`
// (global atomicMin; early-z reads it through L2 when `early_z`).  TILE_MODE = true: keys go to the CTA's
// shared-memory tile `tile_keys` (TILE x TILE, anchored at the item's block origin); early-z always on.
template <bool TILE_MODE, bool BLOCKZ = (RAST_BLOCK_Z != 0)>
RAST_HD void raster_item(const StagedTris &stg, uint32_t it, uint32_t lane, const View &vw, unsigned long long *vis_all,
                                            unsigned long long *tile_keys, bool early_z, const volatile uint32_t *block_far = nullptr, uint32_t live_in = 0xFFFFFFFFu) {
    using namespace exact;
    const uint32_t tri = stg.w[19][it];
    uint32_t live = stg.w[23][it] & live_in; // live_in: the blocks the caller's own test left (k_raster_tiles tests four items per vote)
    if (tri == INVALID_TRI || live == 0u) return;
    const uint32_t qx = (lane & 7u) * 2u, qy = (lane >> 3) * 2u;
    const uint32_t rect0 = stg.w[17][it], rect1 = stg.w[18][it], org = stg.w[24][it];
    const uint32_t rx0 = rect0 & 0xFFFFu, ry0 = rect0 >> 16, rx1 = rect1 & 0xFFFFu, ry1 = rect1 >> 16, ox = org & 0xFFFFu, oy = org >> 16;
#if RAST_BLOCK_Z
    const float bz_o = exact::u2f(stg.w[27][it]), bz_gx = exact::u2f(stg.w[28][it]), bz_gy = exact::u2f(stg.w[29][it]), bz_m = exact::u2f(stg.w[30][it]);
    if (BLOCKZ && TILE_MODE && block_far != nullptr) {
        // Tile schedule, block-level depth rejection for all eight 16 x 8 blocks of the item AT ONCE: lane b tests block b -- the
        // smallest depth the item's plane can take on the block, less the proven margin of stage_item, against the farthest depth
        // stored in that block of the CTA's shared-memory tile (block_far, refreshed by the CTA's warps as they go; a stale value
        // is a larger one, i.e. conservative) -- and one ballot leaves the blocks that still need their edge functions evaluated.
        // At depth complexity 50 five of six blocks lose here; walking the strips and columns only to reject them one by one was
        // two thirds of the kernel's instructions (ncu source counters, 8K overdraw frame).
        const bool lose = lane < 8u && ((live >> lane) & 1u) && block_behind(lane, block_far[lane], rx0, ry0, rx1, ry1, ox, oy, bz_o, bz_gx, bz_gy, bz_m);
        live &= ~RAST_BALLOT(lose);
        if (live == 0u) return;
    }
#endif

    TriSetup s;
    s.x0 = exact::u2f(stg.w[0][it]); s.y0 = exact::u2f(stg.w[1][it]);
    s.x1 = exact::u2f(stg.w[2][it]); s.y1 = exact::u2f(stg.w[3][it]);
    s.x2 = exact::u2f(stg.w[4][it]); s.y2 = exact::u2f(stg.w[5][it]);
    s.z0 = exact::u2f(stg.w[6][it]); s.z1 = exact::u2f(stg.w[7][it]); s.z2 = exact::u2f(stg.w[8][it]);
    s.d12x = exact::u2f(stg.w[9][it]); s.d12y = exact::u2f(stg.w[10][it]);
    s.d20x = exact::u2f(stg.w[11][it]); s.d20y = exact::u2f(stg.w[12][it]);
    s.d01x = exact::u2f(stg.w[13][it]); s.d01y = exact::u2f(stg.w[14][it]);
    s.area = exact::u2f(stg.w[15][it]);
    s.literal = stg.w[16][it] != 0u;
    s.rcp1 = exact::u2f(stg.w[25][it]);
    s.div_ok = stg.w[26][it] != 0u;
    const float rcp_area = exact::u2f(stg.w[21][it]), z_margin = exact::u2f(stg.w[22][it]);
    unsigned long long *vis = TILE_MODE ? nullptr : vis_all + (size_t)stg.w[20][it] * vw.band_pixels;
#pragma unroll 1
    for (uint32_t strip = 0; strip < 4u; ++strip) {
        if (((live >> (strip * 2u)) & 3u) == 0u) continue;
        const uint32_t y = oy + strip * 8u + qy;
        const float pya = (float)y, pyb = (float)(y + 1u);
        // edge k at pixel (i,j): mul(dkx, py_j - yk) - mul(dky, px_i - xk)
        const float a0a = mul(s.d12x, sub(pya, s.y1)), a0b = mul(s.d12x, sub(pyb, s.y1));
        const float a1a = mul(s.d20x, sub(pya, s.y2)), a1b = mul(s.d20x, sub(pyb, s.y2));
        const float a2a = mul(s.d01x, sub(pya, s.y0)), a2b = mul(s.d01x, sub(pyb, s.y0));
        const uint32_t ymask = ((y >= ry0 && y <= ry1) ? 3u : 0u) | ((y + 1u >= ry0 && y + 1u <= ry1) ? 12u : 0u);
#pragma unroll 1
        for (uint32_t column = 0; column < 2u; ++column) {
            if (((live >> (strip * 2u + column)) & 1u) == 0u) continue;
            const uint32_t x = ox + column * 16u + qx;
#if RAST_BLOCK_Z
            // Block-level depth rejection: if the smallest depth the item's plane can take on this block, less the margin, is
            // behind the LARGEST depth stored at the block's pixels (a stored depth only ever decreases, so a stale read is
            // conservative), every fragment of the block would lose its atomicMin: the edge evaluation is skipped for all 128 pixels.
            uint32_t bz_hi[4] = {0u, 0u, 0u, 0u};
            bool bz_loaded = false;
            if (BLOCKZ && !TILE_MODE && early_z) {
                const uint32_t in4 = ymask & (((x >= rx0 && x <= rx1) ? 5u : 0u) | ((x + 1u >= rx0 && x + 1u <= rx1) ? 10u : 0u));
                const unsigned long long *q0 = vis + (size_t)(y - vw.y0) * vw.W + x;
                uint32_t mine = 0u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (in4 & (1u << k)) {
                        bz_hi[k] = RAST_LDCG32(reinterpret_cast<const uint32_t *>(q0 + (k >> 1) * (size_t)vw.W + (k & 1)) + 1);
                        mine = bz_hi[k] > mine ? bz_hi[k] : mine;
                    }
                }
                bz_loaded = true;
                const uint32_t kmax = RAST_WARP_MAX_U32(mine); // empty pixel = 0xFFFFFFFF: nothing is rejected
                if (kmax != 0xFFFFFFFFu) {
                    const uint32_t bxa = max(ox + column * 16u, rx0), bxb = min(ox + column * 16u + 15u, rx1);
                    const uint32_t bya = max(oy + strip * 8u, ry0), byb = min(oy + strip * 8u + 7u, ry1);
                    const float dxa = (float)(bxa - rx0), dxb = (float)(bxb - rx0), dya = (float)(bya - ry0), dyb = (float)(byb - ry0);
#ifndef RAST_BLOCK_Z_TEST_BIAS
#define RAST_BLOCK_Z_TEST_BIAS 0.f // tests/test_emu_device_fns.py builds with a positive bias to show that a wrong bound is caught
#endif
                    const float lb = bz_o + fminf(bz_gx * dxa, bz_gx * dxb) + fminf(bz_gy * dya, bz_gy * dyb) - bz_m + RAST_BLOCK_Z_TEST_BIAS;
                    const uint32_t far_bits = (kmax & 0x80000000u) ? (kmax ^ 0x80000000u) : ~kmax; // inverse of depth_key
                    if (lb > exact::u2f(far_bits)) continue;
                }
            }
#endif
            const float pxa = (float)x, pxb = (float)(x + 1u);
            const float c0a = mul(s.d12y, sub(pxa, s.x1)), c0b = mul(s.d12y, sub(pxb, s.x1));
            const float c1a = mul(s.d20y, sub(pxa, s.x2)), c1b = mul(s.d20y, sub(pxb, s.x2));
            const float c2a = mul(s.d01y, sub(pxa, s.x0)), c2b = mul(s.d01y, sub(pxb, s.x0));
            float e0[4], e1[4], e2[4]; // (xa,ya) (xb,ya) (xa,yb) (xb,yb)
            e0[0] = sub(a0a, c0a); e0[1] = sub(a0a, c0b); e0[2] = sub(a0b, c0a); e0[3] = sub(a0b, c0b);
            e1[0] = sub(a1a, c1a); e1[1] = sub(a1a, c1b); e1[2] = sub(a1b, c1a); e1[3] = sub(a1b, c1b);
            e2[0] = sub(a2a, c2a); e2[1] = sub(a2a, c2b); e2[2] = sub(a2b, c2a); e2[3] = sub(a2b, c2b);
            const uint32_t inrect = ymask & (((x >= rx0 && x <= rx1) ? 5u : 0u) | ((x + 1u >= rx0 && x + 1u <= rx1) ? 10u : 0u));
            uint32_t mask = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (candidate(s, e0[k], e1[k], e2[k])) mask |= 1u << k;
            mask &= inrect;
            if (!RAST_ANY(mask != 0u)) continue;
            // Early depth rejection.  z_est approximates the fragment depth to within z_margin
            // (|z_est - z| <= (2^-22 + 8 ulp) * max|z_k| < z_margin / 2 for a candidate pixel), and a stored depth
            // only ever decreases, so "z_est > stored + margin" proves the exact depth would lose the atomicMin:
            // the divisions and the atomic are skipped.  Stale reads and NaNs fall through to the exact path.
            unsigned long long *pix0 = TILE_MODE ? tile_keys + (strip * 8u + qy) * TILE + column * 16u + qx : vis + (size_t)(y - vw.y0) * vw.W + x;
            const size_t row_stride = TILE_MODE ? (size_t)TILE : (size_t)vw.W;
            if (TILE_MODE || early_z) {
                uint32_t cur_hi[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { // all four reads in flight together (global: L2, bypassing L1)
                    const uint32_t *hi = reinterpret_cast<const uint32_t *>(pix0 + (k >> 1) * row_stride + (k & 1)) + 1;
#if RAST_BLOCK_Z
                    if (bz_loaded) { cur_hi[k] = (mask & (1u << k)) ? bz_hi[k] : 0xFFFFFFFFu; continue; } // read once, for the block test
#endif
                    cur_hi[k] = (mask & (1u << k)) ? (TILE_MODE ? *reinterpret_cast<const volatile uint32_t *>(hi) : RAST_LDCG32(hi)) : 0xFFFFFFFFu;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float z_est = fmaf(s.z2, e2[k], fmaf(s.z1, e1[k], s.z0 * e0[k])) * rcp_area;
                    const uint32_t cur_bits = (cur_hi[k] & 0x80000000u) ? (cur_hi[k] ^ 0x80000000u) : ~cur_hi[k]; // inverse of depth_key
                    if (cur_hi[k] != 0xFFFFFFFFu && z_est > exact::u2f(cur_bits) + z_margin) mask &= ~(1u << k);
                }
                if (!RAST_ANY(mask != 0u)) continue;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mask & (1u << k)) {
                    float b0, b1, b2, z;
                    if (fragment(s, e0[k], e1[k], e2[k], b0, b1, b2, z)) {
                        const unsigned long long key = ((unsigned long long)depth_key(z) << 32) | tri;
                        RAST_ATOMIC_MIN64(pix0 + (k >> 1) * row_stride + (k & 1), key);
                    }
                }
            }
        }
    }
}

// (64 registers / 4 CTAs per SM measured: 1080p Suzanne raster +4 %, 8K overdraw -8 %; the default favours the former)
__global__ void __launch_bounds__(RASTER_WARPS * 32) k_raster_chunks(Scene sc, View vw, Batch bt) {
    __shared__ StagedTris stage_all[RASTER_WARPS];
    if (tile_schedule_taken(bt)) return; // this batch is rasterised by k_raster_tiles
    StagedTris &stg = stage_all[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t count = (uint32_t)min(bt.counters[CNT_QUEUE], (unsigned long long)bt.queue_cap);
    // items per grab: 32 when the queue is long (amortises the fetch latency), fewer when it is short so that
    // a small frame still spreads over the whole grid instead of over count/32 warps
    const uint32_t n_warps = gridDim.x * RASTER_WARPS;
    // (smaller grabs for better balance measured worse at 32 frames per batch -- count / (2 / 4 / 8 n_warps): raster 0.71 / 0.74 / 1.00 against 0.735 ms
    //  per 120 frames -- and the same at 120 per batch, where every warp grabs several times anyway)
    const uint32_t take = max(1u, min(32u, count / n_warps));
    // Early depth rejection pays only when fragments mostly lose: it is switched on when the queued bbox area
    // exceeds EARLY_Z_OVERDRAW times the pixels of the batch (bboxes are about twice the covered area).
    const bool early_z = bt.counters[CNT_BBOX_AREA] > (unsigned long long)EARLY_Z_OVERDRAW * vw.band_pixels * bt.n_frames;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = (uint32_t)min(atomicAdd(&bt.counters[CNT_CURSOR], (unsigned long long)take), 0xFFFFFFFFull);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= count) break;
        const uint32_t n_items = min(take, count - base);

        {   // ---- stage: lane = item ----
            uint32_t tri = INVALID_TRI, f = 0, rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0;
            if (lane < n_items) {
                const uint2 item = bt.queue[base + lane];
                tri = item.x;
                if (tri != INVALID_TRI) {
                    const uint32_t cx = item.y & 0xFFFu, cy = (item.y >> 12) & 0xFFFu;
                    f = item.y >> 24;
                    const float4 *rv = bt.rv + (size_t)f * sc.V;
                    v0 = rv[sc.vidx0[tri]]; v1 = rv[sc.vidx1[tri]]; v2 = rv[sc.vidx2[tri]];
                    const BBox bb = bounding_box(v0, v1, v2, vw);
                    rx0 = bb.x0 + cx * CHUNK; ry0 = bb.y0 + cy * CHUNK;
                    rx1 = min(bb.x1, rx0 + CHUNK - 1u); ry1 = min(bb.y1, ry0 + CHUNK - 1u);
                }
            }
            stage_item(stg, lane, tri, f, v0, v1, v2, rx0, ry0, rx1, ry1, rx0, ry0);
            if (tri != INVALID_TRI && stg.w[23][lane] != 0u) mark_tiles(bt.tile_flags, f, vw, rx0, ry0, rx1, ry1); // some block of the item may hold a candidate
        }
        __syncwarp();
        if (RAST_BLOCK_Z != 0 && early_z) {
            for (uint32_t it = 0; it < n_items; ++it) raster_item<false, true>(stg, it, lane, vw, bt.vis, nullptr, true);
        } else {
            for (uint32_t it = 0; it < n_items; ++it) raster_item<false, false>(stg, it, lane, vw, bt.vis, nullptr, early_z);
        }
        __syncwarp();
    }
}

// ---- screen-tile binning ----------------------------------------------------------------------
// (Round 2 began with three passes -- k_setup counted per tile, a single-CTA k_plan_tiles scanned the counts, k_fill_tiles scattered the
// triangles into exact runs: 86 + 94 us for the 32 400 tiles / 200 k triangles of the 8K overdraw frame, 10 % of it.  Bins of fixed
// capacity filled by k_setup itself need neither.)
// One CTA per (tile, frame): the tile's 1024 keys live in shared memory until every triangle of the bin is done, so the depth
// test, the early rejections and the atomics never leave the SM (the chunk queue re-read 10.8 GB of keys from DRAM on the 8K
// overdraw frame: 265 MB of keys, 2 x the L2, visited in random order).  What makes the schedule pay at depth complexity 50:
//   * the bin is processed NEAR TO FAR: a counting sort of its entries by the depth key of the triangle's nearest vertex (64
//     buckets between the bin's own minimum and maximum, in shared memory; bins above TILE_SORT_MAX entries stay unsorted --
//     the order only decides how much work is skipped, never the result);
//   * every warp keeps refreshing the FARTHEST stored depth of "its" two 16 x 8 blocks (block_far); an item whose depth plane,
//     less the proven margin of stage_item, lies behind it on a block skips that block without evaluating an edge function,
//     and an item that loses on all its blocks is dropped when it is staged.  After the first few layers almost everything is;
//   * 128 items are staged at a time by the whole CTA (lane = item) and then dealt to the four warps round-robin, so the near
//     items -- the ones that do real work -- are spread over all warps.
// Eight CTAs per SM: 64 registers (__launch_bounds__) and at most 28 032 bytes of shared memory each -- the sorted order is kept as 16-bit
// positions in the bin, 1792 of them (measured on the 8K overdraw frame: 6 CTAs / 80 registers 1.738 ms, 7 / 72: 1.654, 8 / 64: 1.615).
constexpr uint32_t TILE_WARPS = 4, TILE_SORT_MAX = 1792, TILE_SORT_BUCKETS = 64;

__device__ __forceinline__ float key_to_depth(uint32_t k) { return exact::u2f((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); } // inverse of depth_key

#ifndef RAST_TILE_MIN_BLOCKS
#define RAST_TILE_MIN_BLOCKS 8
#endif
#if RAST_TILE_MIN_BLOCKS > 0
__global__ void __launch_bounds__(TILE_WARPS * 32, RAST_TILE_MIN_BLOCKS) k_raster_tiles(Scene sc, View vw, Batch bt, TileBins tb) {
#else
__global__ void __launch_bounds__(TILE_WARPS * 32) k_raster_tiles(Scene sc, View vw, Batch bt, TileBins tb) {
#endif
    __shared__ StagedTris stage_all[TILE_WARPS];
    __shared__ unsigned long long tile_keys[TILE * TILE];
    __shared__ uint16_t order[TILE_SORT_MAX];           // position in the bin of the k-th nearest item
    __shared__ uint32_t hist[TILE_SORT_BUCKETS];
    __shared__ uint32_t block_far[8];
    __shared__ uint32_t zrange[2];
    if (!tile_schedule_taken(bt)) return;
    const uint32_t f = blockIdx.y, tile = blockIdx.x;
    const uint32_t g = f * tb.tiles_x * tb.tiles_y + tile;
    if (g == 0u && threadIdx.x == 0u) bt.counters[CNT_TILE_MODE] = 1ull; // for the host: the bins were used (rast_last_schedule, statistics)
    const uint32_t n = min(tb.fill[g], tb.cap);
    if (n == 0u) return;
    const uint32_t ox = (tile % tb.tiles_x) * TILE, oy = vw.y0 + (tile / tb.tiles_x) * TILE;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // pixels of the tile that lie outside the image can never be written: key 0 (nearer than anything) keeps them out of the far bounds
    for (uint32_t i = tid; i < TILE * TILE; i += TILE_WARPS * 32) tile_keys[i] = (ox + (i % TILE) < vw.W && oy + (i / TILE) < vw.y1) ? VIS_EMPTY : 0ull;
    if (tid < 8u) block_far[tid] = 0xFFFFFFFFu;
    if (tid < TILE_SORT_BUCKETS) hist[tid] = 0u;
    if (tid == 0u) { zrange[0] = 0xFFFFFFFFu; zrange[1] = 0u; }
    __syncthreads();

    // ---- near-to-far order of the bin (counting sort on the nearest-vertex depth key) ----
    const uint2 *bin = tb.items + (size_t)g * tb.cap;
    const bool sorted = n <= TILE_SORT_MAX;
    if (sorted) {
        uint32_t lo = 0xFFFFFFFFu, hi = 0u;
        for (uint32_t i = tid; i < n; i += TILE_WARPS * 32) { const uint32_t z = bin[i].y; lo = min(lo, z); hi = max(hi, z); }
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (lane == 0u) { atomicMin(&zrange[0], lo); atomicMax(&zrange[1], hi); }
        __syncthreads();
        const uint32_t zlo = zrange[0];
        const float scale = (float)(TILE_SORT_BUCKETS - 1u) / fmaxf((float)(zrange[1] - zlo), 1.0f); // monotone in the key: equal keys share a bucket
        for (uint32_t i = tid; i < n; i += TILE_WARPS * 32) atomicAdd(&hist[min((uint32_t)((float)(bin[i].y - zlo) * scale), TILE_SORT_BUCKETS - 1u)], 1u);
        __syncthreads();
        if (warp == 0u) { // exclusive scan of the 64 bucket counts: two per lane
            const uint32_t a = hist[2u * lane], b = hist[2u * lane + 1u];
            uint32_t incl = a + b;
#pragma unroll
            for (uint32_t d = 1; d < 32u; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
            hist[2u * lane] = incl - a - b;
            hist[2u * lane + 1u] = incl - b;
        }
        __syncthreads();
        for (uint32_t i = tid; i < n; i += TILE_WARPS * 32) {
            const uint2 e = bin[i];
            order[atomicAdd(&hist[min((uint32_t)((float)(e.y - zlo) * scale), TILE_SORT_BUCKETS - 1u)], 1u)] = (uint16_t)i;
        }
        __syncthreads();
    }

    const float4 *rv = bt.rv + (size_t)f * sc.V;
    // the farthest stored depth of block b of the tile (strip b >> 1, column b & 1): lane -> row lane >> 2, four adjacent pixels
    auto refresh_far = [&](uint32_t b) {
        const volatile unsigned long long *p = tile_keys + ((b >> 1) * 8u + (lane >> 2)) * TILE + (b & 1u) * 16u + (lane & 3u) * 4u;
        uint32_t m = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) m = max(m, (uint32_t)(p[k] >> 32));
        m = __reduce_max_sync(0xFFFFFFFFu, m);
        if (lane == 0u) *reinterpret_cast<volatile uint32_t *>(&block_far[b]) = m;
    };
    for (uint32_t base = 0; base < n; base += TILE_WARPS * 32) {
        const uint32_t m_items = min(TILE_WARPS * 32u, n - base);
        {   // ---- stage: thread = item ----
            StagedTris &stg = stage_all[warp];
            uint32_t tri = INVALID_TRI, rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0;
            if (tid < m_items) {
                tri = bin[sorted ? (uint32_t)order[base + tid] : base + tid].x;
                v0 = rv[sc.vidx0[tri]]; v1 = rv[sc.vidx1[tri]]; v2 = rv[sc.vidx2[tri]];
                const BBox bb = bounding_box(v0, v1, v2, vw);
                rx0 = max(bb.x0, ox); ry0 = max(bb.y0, oy);
                rx1 = min(bb.x1, ox + TILE - 1u); ry1 = min(bb.y1, oy + TILE - 1u);
                if (rx0 > rx1 || ry0 > ry1) tri = INVALID_TRI;
            }
            stage_item(stg, lane, tri, f, v0, v1, v2, rx0, ry0, rx1, ry1, ox, oy);
#if RAST_BLOCK_Z
            if (tri != INVALID_TRI && stg.w[23][lane] != 0u) {
                // the whole item against the farthest depth stored anywhere in the tile: lower bound of its depth plane over its rectangle
                uint32_t kmax = 0u;
#pragma unroll
                for (int b = 0; b < 8; ++b) kmax = max(kmax, *reinterpret_cast<volatile uint32_t *>(&block_far[b]));
                if (kmax != 0xFFFFFFFFu) {
                    const float gx = exact::u2f(stg.w[28][lane]), gy = exact::u2f(stg.w[29][lane]);
                    const float lb = exact::u2f(stg.w[27][lane]) + fminf(0.f, gx * (float)(rx1 - rx0)) + fminf(0.f, gy * (float)(ry1 - ry0)) - exact::u2f(stg.w[30][lane]);
                    if (lb > key_to_depth(kmax)) stg.w[23][lane] = 0u; // behind everything already in the tile: no live block
                }
            }
#endif
        }
        __syncthreads();
        // ---- rasterise: item j of the round goes to warp j % 4 (the nearest items first, one per warp) ----
        // (compacting the items that still have a live block -- half of a bin's have none -- into a list before dealing them out was
        //  measured: 1.757 vs 1.738 ms on the 8K overdraw frame, the extra barrier costs what the skipped visits save)
#if RAST_BLOCK_Z
        // Four of the warp's items per vote: lanes 8q .. 8q + 7 test the eight blocks of the warp's q-th next item against block_far (the
        // test of raster_item, a quarter of the warp per item instead of the whole warp per item -- at depth complexity 50 most visits end
        // right there), and only items that keep a block are rasterised.  Items 1 .. 3 of a group are tested before item 0 has written
        // its keys (a stale farthest depth is a larger one, i.e. conservative), so raster_item repeats the test for the survivors with the
        // fresh values: without that 28 % more items and 26 % more blocks reached the edge functions (ncu).
        for (uint32_t j0 = warp; j0 < m_items; j0 += TILE_WARPS * 4u) {
            refresh_far(2u * warp); refresh_far(2u * warp + 1u);
            const uint32_t j = j0 + (lane >> 3) * TILE_WARPS, b = lane & 7u;
            bool keep = false;
            if (j < m_items) {
                const StagedTris &sg = stage_all[j >> 5];
                const uint32_t it = j & 31u;
                if (sg.w[19][it] != INVALID_TRI && ((sg.w[23][it] >> b) & 1u)) {
                    const uint32_t rect0 = sg.w[17][it], rect1 = sg.w[18][it], org = sg.w[24][it];
                    keep = !block_behind(b, *reinterpret_cast<volatile uint32_t *>(&block_far[b]), rect0 & 0xFFFFu, rect0 >> 16, rect1 & 0xFFFFu, rect1 >> 16, org & 0xFFFFu, org >> 16,
                                         exact::u2f(sg.w[27][it]), exact::u2f(sg.w[28][it]), exact::u2f(sg.w[29][it]), exact::u2f(sg.w[30][it]));
                }
            }
            const uint32_t masks = __ballot_sync(0xFFFFFFFFu, keep);
#pragma unroll 1
            for (uint32_t q = 0; q < 4u; ++q) {
                const uint32_t m = (masks >> (8u * q)) & 0xFFu, jj = j0 + q * TILE_WARPS;
                if (m != 0u) raster_item<true, true>(stage_all[jj >> 5], jj & 31u, lane, vw, nullptr, tile_keys, true, block_far, m); // (tests its blocks once more, freshly)
            }
        }
#else
        uint32_t since = 0;
        for (uint32_t j = warp; j < m_items; j += TILE_WARPS, ++since) {
            if ((since & 3u) == 0u) { refresh_far(2u * warp); refresh_far(2u * warp + 1u); }
            raster_item<true, false>(stage_all[j >> 5], j & 31u, lane, vw, nullptr, tile_keys, true);
        }
#endif
        refresh_far(2u * warp); refresh_far(2u * warp + 1u);
        __syncthreads();
    }
    // merge the tile into the visibility buffer (tiny triangles were written there directly by k_setup)
    unsigned long long *vis = bt.vis + (size_t)f * vw.band_pixels;
    bool any = false;
    for (uint32_t i = tid; i < TILE * TILE; i += TILE_WARPS * 32) {
        const unsigned long long key = tile_keys[i];
        const uint32_t x = ox + (i % TILE), y = oy + (i / TILE);
        if (key != VIS_EMPTY && x < vw.W && y < vw.y1) {
            atomicMin(vis + (size_t)(y - vw.y0) * vw.W + x, key);
            any = true;
        }
    }
    // thread tid covers columns tid % 32 of rows tid / 32 + 4k: warp w = rows w, w + 4, ...: both tile-flag rows (16 pixel rows each)
    if (__any_sync(0xFFFFFFFFu, any) && lane == 0u) mark_tiles(bt.tile_flags, f, vw, ox, oy, min(ox + TILE - 1u, vw.W - 1u), min(oy + TILE - 1u, vw.y1 - 1u));
}

// ---- K4: resolve + deferred shading ---------------------------------------------------------
// Material::sample on a textured material (material.cpp:19-21): three CImg::_linear_atXY lookups
// (CImg.h:13475-13492) at the same position.  The position arithmetic is shared between the channels and
// the texels are stored interleaved (r, g, b, -) on the device, so the four corners are four 16-byte loads
// instead of twelve 4-byte ones; the per-channel arithmetic is CImg's, unchanged.
RAST_HD void sample_texture(const float4 *__restrict__ tex, int w, int h, float fx, float fy, float &r, float &g, float &b) {
    using namespace exact;
    const float hx = (float)(w - 1), hy = (float)(h - 1);
    const float nfx = fx < 0.f ? 0.f : (fx > hx ? hx : fx); // cimg::cut (CImg.h:5184-5186)
    const float nfy = fy < 0.f ? 0.f : (fy > hy ? hy : fy);
    const uint32_t x = to_uint(nfx), y = to_uint(nfy);
    const float dx = sub(nfx, (float)x), dy = sub(nfy, (float)y);
    const uint32_t nx = dx > 0.f ? x + 1u : x, ny = dy > 0.f ? y + 1u : y;
    const float4 cc = exact::ldg(tex + (x + y * (uint32_t)w)), nc = exact::ldg(tex + (nx + y * (uint32_t)w));
    const float4 cn = exact::ldg(tex + (x + ny * (uint32_t)w)), nn = exact::ldg(tex + (nx + ny * (uint32_t)w));
    // Icc + dx*(Inc - Icc + dy*(Icc + Inn - Icn - Inc)) + dy*(Icn - Icc)
    r = add(add(cc.x, mul(dx, add(sub(nc.x, cc.x), mul(dy, sub(sub(add(cc.x, nn.x), cn.x), nc.x))))), mul(dy, sub(cn.x, cc.x)));
    g = add(add(cc.y, mul(dx, add(sub(nc.y, cc.y), mul(dy, sub(sub(add(cc.y, nn.y), cn.y), nc.y))))), mul(dy, sub(cn.y, cc.y)));
    b = add(add(cc.z, mul(dx, add(sub(nc.z, cc.z), mul(dy, sub(sub(add(cc.z, nn.z), cn.z), nc.z))))), mul(dy, sub(cn.z, cc.z)));
}

struct Shaded { uint32_t r, g, b; float depth; };

constexpr uint32_t PARAM_LIGHTS = 64;

// The first PARAM_LIGHTS lights travel as a kernel parameter (constant bank: uniform loads, no staging, no
// barrier): (-trans_dir.xyz, intensity*colour.r) (intensity*colour.g, intensity*colour.b).
struct LightTable {
    uint32_t n;          // all lights
    uint32_t pad[3];
    float4 a[PARAM_LIGHTS];
    float2 c[PARAM_LIGHTS];
};

// The shading half of update_pixel (drawing.cpp:121-146) for the winning triangle of one pixel.
template <bool PRE_NORMALS, bool FLAT>
RAST_HD Shaded shade_pixel(uint32_t tri, uint32_t x, uint32_t y, const Scene &sc, const float4 *__restrict__ rv,
                                              const float4 *__restrict__ cn, const float *__restrict__ normal_m, bool wind_clockwise,
                                              const LightTable &lt, const LightDev *__restrict__ lights) {
    using namespace exact;
    Shaded out;
    const uint4 *rec = reinterpret_cast<const uint4 *>(sc.tri_rec) + 3 * (size_t)tri; // indices are non-negative after upload
    const uint4 r0 = exact::ldg(rec), r1 = exact::ldg(rec + 1);
    const uint2 r2 = exact::ldg(reinterpret_cast<const uint2 *>(rec + 2));
    const float4 v0 = exact::ldg(rv + r0.x), v1 = exact::ldg(rv + r0.y), v2 = exact::ldg(rv + r0.z); // written by k_vertex, read-only here

    // vertex normals in camera space (transform_direction, geometry.cpp:35-42,97-108)
    float4 n0, n1, n2;
    float fx = 0.f, fy = 0.f, fz = 0.f; // FLAT: un-normalised face normal
    if (FLAT) {
        // extension: cross(c1 - c0, c2 - c0) of the camera-space vertices xyz(modelview * (v, 1)) (drawing.cpp:232-233);
        // normal_m points at the frame's modelview matrix in this mode
        float mv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) mv[k] = exact::ldg(normal_m + k);
        const float *p0 = sc.pos + 3 * (size_t)r0.x, *p1 = sc.pos + 3 * (size_t)r0.y, *p2 = sc.pos + 3 * (size_t)r0.z;
        const float4 c0 = mat_vec(mv, exact::ldg(p0), exact::ldg(p0 + 1), exact::ldg(p0 + 2), 1.f);
        const float4 c1 = mat_vec(mv, exact::ldg(p1), exact::ldg(p1 + 1), exact::ldg(p1 + 2), 1.f);
        const float4 c2 = mat_vec(mv, exact::ldg(p2), exact::ldg(p2 + 1), exact::ldg(p2 + 2), 1.f);
        const float ax = sub(c1.x, c0.x), ay = sub(c1.y, c0.y), az = sub(c1.z, c0.z);
        const float bx = sub(c2.x, c0.x), by = sub(c2.y, c0.y), bz = sub(c2.z, c0.z);
        fx = sub(mul(ay, bz), mul(by, az)); // glm::cross
        fy = sub(mul(az, bx), mul(bz, ax));
        fz = sub(mul(ax, by), mul(bx, ay));
        n0 = n1 = n2 = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (PRE_NORMALS) {
        n0 = exact::ldg(cn + r0.w); n1 = exact::ldg(cn + r1.x); n2 = exact::ldg(cn + r1.y);
    } else {
        const float4 m0 = exact::ldg(sc.nrm4 + r0.w), m1 = exact::ldg(sc.nrm4 + r1.x), m2 = exact::ldg(sc.nrm4 + r1.y);
        float nm[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) nm[k] = exact::ldg(normal_m + k);
        n0 = mat_vec(nm, m0.x, m0.y, m0.z, 0.f);
        n1 = mat_vec(nm, m1.x, m1.y, m1.z, 0.f);
        n2 = mat_vec(nm, m2.x, m2.y, m2.z, 0.f);
    }
    const MaterialDev *mp = sc.mats + r2.y;
    const float4 mk = exact::ldg(reinterpret_cast<const float4 *>(mp));      // kd.rgb, has_texture
    const int4 mt = exact::ldg(reinterpret_cast<const int4 *>(mp) + 1);      // tex_w, tex_h, texel_offset (lo, hi)

    // barycentric + depth, the same operations as the raster pass => the same bits (drawing.cpp:41-49,115-116)
    const float px = (float)x, py = (float)y;
    const float area = sub(mul(sub(v1.x, v0.x), sub(v2.y, v0.y)), mul(sub(v1.y, v0.y), sub(v2.x, v0.x)));
    const float e0 = sub(mul(sub(v2.x, v1.x), sub(py, v1.y)), mul(sub(v2.y, v1.y), sub(px, v1.x)));
    const float e1 = sub(mul(sub(v0.x, v2.x), sub(py, v2.y)), mul(sub(v0.y, v2.y), sub(px, v2.x)));
    const float e2 = sub(mul(sub(v1.x, v0.x), sub(py, v0.y)), mul(sub(v1.y, v0.y), sub(px, v0.x)));
    float b0, b1, b2;
    div3(e0, e1, e2, area, div_reciprocal(area), div_in_range(area), b0, b1, b2);
    out.depth = add(add(mul(v0.z, b0), mul(v1.z, b1)), mul(v2.z, b2));

    // interpolation_coords + camera-space depth (drawing.cpp:125-128)
    const float i0 = mul(v0.w, b0), i1 = mul(v1.w, b1), i2 = mul(v2.w, b2);
    const float d = rcp(add(add(i0, i1), i2));

    // Material::sample (material.cpp:11-26).  The texel loads are issued before the normal is interpolated and normalised (a
    // division and a square root, ~80 instructions) so that their latency is covered by this warp's own arithmetic
    // (8.20 -> 8.14 ms per 720 frames on the 1080p spin).
    float ar = mk.x, ag = mk.y, ab = mk.z;
    if (exact::f2u(mk.w) != 0u) {
        const float2 uv0 = exact::ldg(sc.uv + r1.z), uv1 = exact::ldg(sc.uv + r1.w), uv2 = exact::ldg(sc.uv + r2.x);
        const float u = mul(d, add(add(mul(i0, uv0.x), mul(i1, uv1.x)), mul(i2, uv2.x))); // drawing.cpp:135
        const float v = mul(d, add(add(mul(i0, uv0.y), mul(i1, uv1.y)), mul(i2, uv2.y)));
        const long long toff = ((long long)(uint32_t)mt.z) | ((long long)mt.w << 32);
        unsigned long long tex_base = (unsigned long long)(sc.texels + toff);
        asm volatile("" : "+l"(tex_base)); // opaque base: each corner becomes one IMAD.WIDE instead of a 64-bit add + LEA pair
        sample_texture(reinterpret_cast<const float4 *>(tex_base), mt.x, mt.y, mul(u, (float)mt.x), mul(sub(1.f, v), (float)mt.y), ar, ag, ab);
        if (exact::f2u(mk.w) & 2u) { ar = mul(mk.x, ar); ag = mul(mk.y, ag); ab = mul(mk.z, ab); } // extension RAST_TEXTURE_MODULATE_KD: texel x Kd
    }

    // perspective_interpolate + normalize (drawing.cpp:64-75,131-132)
    const float mx = FLAT ? fx : mul(d, add(add(mul(i0, n0.x), mul(i1, n1.x)), mul(i2, n2.x)));
    const float my = FLAT ? fy : mul(d, add(add(mul(i0, n0.y), mul(i1, n1.y)), mul(i2, n2.y)));
    const float mz = FLAT ? fz : mul(d, add(add(mul(i0, n0.z), mul(i1, n1.z)), mul(i2, n2.z)));
    const float inv = rcp(fsqrt(add(add(mul(mx, mx), mul(my, my)), mul(mz, mz))));
    float nx = mul(mx, inv), ny = mul(my, inv), nz = mul(mz, inv);
    if (wind_clockwise) { nx = -nx; ny = -ny; nz = -nz; }

    // shade / light_contribution (shading.cpp:20-34)
    float sr = 0.f, sg = 0.f, sb = 0.f;
    const uint32_t n_p = lt.n < PARAM_LIGHTS ? lt.n : PARAM_LIGHTS;
    auto one_light = [&](uint32_t l) {
        const float4 a = lt.a[l];
        const float2 c = lt.c[l];
        const float k = glm_max(0.f, add(add(mul(nx, a.x), mul(ny, a.y)), mul(nz, a.z)));
        sr = add(sr, mul(mul(mul(a.w, ar), k), 0.318309886183790671537767526745028724f));
        sg = add(sg, mul(mul(mul(c.x, ag), k), 0.318309886183790671537767526745028724f));
        sb = add(sb, mul(mul(mul(c.y, ab), k), 0.318309886183790671537767526745028724f));
    };
    // the light count is uniform over the launch: the common small counts run straight-line code with the
    // light constants at immediate constant-bank offsets (same lights, same order, same operations)
    if (n_p == 3u) {
        one_light(0); one_light(1); one_light(2);
    } else if (n_p == 1u) {
        one_light(0);
    } else if (n_p == 2u) {
        one_light(0); one_light(1);
    } else if (n_p == 4u) {
        one_light(0); one_light(1); one_light(2); one_light(3);
    } else {
        uint32_t l = 0;
#pragma unroll 1
        for (; l + 4u <= n_p; l += 4u) { one_light(l); one_light(l + 1u); one_light(l + 2u); one_light(l + 3u); }
#pragma unroll 1
        for (; l < n_p; ++l) one_light(l);
    }
#pragma unroll 1
    for (uint32_t l = n_p; l < lt.n; ++l) { // beyond the parameter table: straight from global memory
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

#if RAST_SHADE_PREP
// ---- prepared shading records (variant) -----------------------------------------------------------------------------
// q0 (v0.x v0.y v1.x v1.y)  q1 (v2.x v2.y area rcp1)  q2 (d12x d12y d20x d20y)  q3 (d01x d01y div_ok material)
// q4 (z0 z1 z2 w0)  q5 (w1 w2 uv0.x uv0.y)  q6 (uv1.x uv1.y uv2.x uv2.y)  q7 (n0.x n0.y n0.z n1.x)  q8 (n1.y n1.z n2.x n2.y)  q9 (n2.z - - -)
// with dij = vj - vi exactly as shade_pixel forms them, area = edge(v2; v0, v1) (drawing.cpp:46), rcp1 / div_ok as
// exact::div3 wants them.  Everything in it is a value shade_pixel would compute or load for every pixel of the triangle.
constexpr uint32_t PREP_QUADS = 10;

RAST_HD void prepare_triangle(uint32_t tri, const Scene &sc, const float4 *__restrict__ rv, const float4 *__restrict__ cn, float4 *__restrict__ out) {
    using namespace exact;
    const uint4 *rec = reinterpret_cast<const uint4 *>(sc.tri_rec) + 3 * (size_t)tri;
    const uint4 r0 = ldg(rec), r1 = ldg(rec + 1);
    const uint2 r2 = ldg(reinterpret_cast<const uint2 *>(rec + 2));
    const float4 v0 = rv[r0.x], v1 = rv[r0.y], v2 = rv[r0.z];
    const float4 n0 = cn[r0.w], n1 = cn[r1.x], n2 = cn[r1.y];
    const float2 uv0 = ldg(sc.uv + r1.z), uv1 = ldg(sc.uv + r1.w), uv2 = ldg(sc.uv + r2.x);
    const float d01x = sub(v1.x, v0.x), d01y = sub(v1.y, v0.y);
    const float area = sub(mul(d01x, sub(v2.y, v0.y)), mul(d01y, sub(v2.x, v0.x)));
    out[0] = make_float4(v0.x, v0.y, v1.x, v1.y);
    out[1] = make_float4(v2.x, v2.y, area, div_reciprocal(area));
    out[2] = make_float4(sub(v2.x, v1.x), sub(v2.y, v1.y), sub(v0.x, v2.x), sub(v0.y, v2.y));
    out[3] = make_float4(d01x, d01y, u2f(div_in_range(area) ? 1u : 0u), u2f(r2.y));
    out[4] = make_float4(v0.z, v1.z, v2.z, v0.w);
    out[5] = make_float4(v1.w, v2.w, uv0.x, uv0.y);
    out[6] = make_float4(uv1.x, uv1.y, uv2.x, uv2.y);
    out[7] = make_float4(n0.x, n0.y, n0.z, n1.x);
    out[8] = make_float4(n1.y, n1.z, n2.x, n2.y);
    out[9] = make_float4(n2.z, 0.f, 0.f, 0.f);
}

// one thread per (triangle, frame); back faces are prepared too (968 x 32 records per batch on the spin sequence: nothing)
__global__ void __launch_bounds__(128) k_prepare_tris(Scene sc, Batch bt) {
    const uint64_t t = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const uint32_t f = blockIdx.y;
    if (t >= sc.T) return;
    prepare_triangle((uint32_t)t, sc, bt.rv + (size_t)f * sc.V, bt.cn + (size_t)f * sc.Nn, bt.prep + ((size_t)f * sc.T + t) * PREP_QUADS);
}

// A prepared record and its material in registers, loaded per covered pixel.  (Keeping it per lane across the rows of a tile and
// reloading only when the lane's winning triangle changes was measured on a B200 -- a lane walks down a pixel column and a Suzanne
// triangle is ~50 pixels tall, so ten of the twelve 16-byte loads per pixel disappear: L1 data-pipe utilisation 73 % -> 45 %, but 80
// instead of 64 registers, 31 % instead of 40 % of the warp slots occupied, and the 32-frame 1080p launch went from 339 to 387 us.
// The pass is co-limited by issue slots and the L1 data pipe; trading one for occupancy does not pay.)
struct PrepRec {
    float4 q0, q1, q2, q3, q4, q5, q6, q7, q8;
    float n2z;
    float4 mk;               // material: kd.rgb, has_texture bits
    int tex_w, tex_h;
    unsigned long long tex_base;
};

RAST_HD void load_prep_record(PrepRec &r, uint32_t tri, const Scene &sc, const float4 *__restrict__ prep) {
    using namespace exact;
    const float4 *q = prep + (size_t)tri * PREP_QUADS;
    r.q0 = ldg(q); r.q1 = ldg(q + 1); r.q2 = ldg(q + 2); r.q3 = ldg(q + 3); r.q4 = ldg(q + 4); r.q5 = ldg(q + 5);
    r.q6 = ldg(q + 6); r.q7 = ldg(q + 7); r.q8 = ldg(q + 8);
    r.n2z = ldg(reinterpret_cast<const float *>(q + 9));
    const MaterialDev *mp = sc.mats + f2u(r.q3.w);
    r.mk = ldg(reinterpret_cast<const float4 *>(mp));                   // kd.rgb, has_texture
    const int4 mt = ldg(reinterpret_cast<const int4 *>(mp) + 1);        // tex_w, tex_h, texel_offset (lo, hi)
    r.tex_w = mt.x; r.tex_h = mt.y;
    const long long toff = ((long long)(uint32_t)mt.z) | ((long long)mt.w << 32);
    r.tex_base = (unsigned long long)(sc.texels + toff);
#ifdef __CUDA_ARCH__
    asm volatile("" : "+l"(r.tex_base)); // opaque base: each corner becomes one IMAD.WIDE instead of a 64-bit add + LEA pair
#endif
}

// shade_pixel for a prepared triangle: the same operations on the same values, in the same order, from the record
RAST_HD Shaded shade_prepared(const PrepRec &p, uint32_t x, uint32_t y, bool wind_clockwise, const LightTable &lt, const LightDev *__restrict__ lights) {
    using namespace exact;
    Shaded out;
    const float4 q0 = p.q0, q1 = p.q1, q2 = p.q2, q3 = p.q3, q4 = p.q4, q5 = p.q5, q7 = p.q7, q8 = p.q8;
    const float n2z = p.n2z;
    const float4 mk = p.mk;

    // barycentric + depth (drawing.cpp:41-49,115-116): e_k from the prepared differences
    const float px = (float)x, py = (float)y;
    const float e0 = sub(mul(q2.x, sub(py, q0.w)), mul(q2.y, sub(px, q0.z)));
    const float e1 = sub(mul(q2.z, sub(py, q1.y)), mul(q2.w, sub(px, q1.x)));
    const float e2 = sub(mul(q3.x, sub(py, q0.y)), mul(q3.y, sub(px, q0.x)));
    float b0, b1, b2;
    div3(e0, e1, e2, q1.z, q1.w, f2u(q3.z) != 0u, b0, b1, b2);
    out.depth = add(add(mul(q4.x, b0), mul(q4.y, b1)), mul(q4.z, b2));

    const float i0 = mul(q4.w, b0), i1 = mul(q5.x, b1), i2 = mul(q5.y, b2);
    const float d = rcp(add(add(i0, i1), i2));

    float ar = mk.x, ag = mk.y, ab = mk.z;
    if (f2u(mk.w) != 0u) {
        const float4 q6 = p.q6;
        const float u = mul(d, add(add(mul(i0, q5.z), mul(i1, q6.x)), mul(i2, q6.z))); // drawing.cpp:135
        const float v = mul(d, add(add(mul(i0, q5.w), mul(i1, q6.y)), mul(i2, q6.w)));
        sample_texture(reinterpret_cast<const float4 *>(p.tex_base), p.tex_w, p.tex_h, mul(u, (float)p.tex_w), mul(sub(1.f, v), (float)p.tex_h), ar, ag, ab);
        if (exact::f2u(mk.w) & 2u) { ar = mul(mk.x, ar); ag = mul(mk.y, ag); ab = mul(mk.z, ab); } // extension RAST_TEXTURE_MODULATE_KD: texel x Kd
    }

    const float mx = mul(d, add(add(mul(i0, q7.x), mul(i1, q7.w)), mul(i2, q8.z)));
    const float my = mul(d, add(add(mul(i0, q7.y), mul(i1, q8.x)), mul(i2, q8.w)));
    const float mz = mul(d, add(add(mul(i0, q7.z), mul(i1, q8.y)), mul(i2, n2z)));
    const float inv = rcp(fsqrt(add(add(mul(mx, mx), mul(my, my)), mul(mz, mz))));
    float nx = mul(mx, inv), ny = mul(my, inv), nz = mul(mz, inv);
    if (wind_clockwise) { nx = -nx; ny = -ny; nz = -nz; }

    float sr = 0.f, sg = 0.f, sb = 0.f;
    const uint32_t n_p = lt.n < PARAM_LIGHTS ? lt.n : PARAM_LIGHTS;
    auto one_light = [&](uint32_t l) {
        const float4 a = lt.a[l];
        const float2 c = lt.c[l];
        const float k = glm_max(0.f, add(add(mul(nx, a.x), mul(ny, a.y)), mul(nz, a.z)));
        sr = add(sr, mul(mul(mul(a.w, ar), k), 0.318309886183790671537767526745028724f));
        sg = add(sg, mul(mul(mul(c.x, ag), k), 0.318309886183790671537767526745028724f));
        sb = add(sb, mul(mul(mul(c.y, ab), k), 0.318309886183790671537767526745028724f));
    };
    if (n_p == 3u) {
        one_light(0); one_light(1); one_light(2);
    } else if (n_p == 1u) {
        one_light(0);
    } else if (n_p == 2u) {
        one_light(0); one_light(1);
    } else if (n_p == 4u) {
        one_light(0); one_light(1); one_light(2); one_light(3);
    } else {
        uint32_t l = 0;
#pragma unroll 1
        for (; l + 4u <= n_p; l += 4u) { one_light(l); one_light(l + 1u); one_light(l + 2u); one_light(l + 3u); }
#pragma unroll 1
        for (; l < n_p; ++l) one_light(l);
    }
#pragma unroll 1
    for (uint32_t l = n_p; l < lt.n; ++l) {
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

// one pixel, record loaded on the spot (host emulation, tests)
RAST_HD Shaded shade_pixel_prep(uint32_t tri, uint32_t x, uint32_t y, const Scene &sc, const float4 *__restrict__ prep, bool wind_clockwise,
                                const LightTable &lt, const LightDev *__restrict__ lights) {
    PrepRec r;
    load_prep_record(r, tri, sc, prep);
    return shade_prepared(r, x, y, wind_clockwise, lt, lights);
}
#endif // RAST_SHADE_PREP

// `covered` bit r of every lane: this lane's pixel (column x0 + lane) of the warp's row r has a winning triangle.  row_spans points at the
// (min x, min W-1-x) pair of the warp's first row.
template <uint32_t ROWS>
__device__ __forceinline__ void note_row_spans(uint32_t *row_spans, uint32_t covered, uint32_t nrows, uint32_t lane, uint32_t x0, uint32_t W) {
    uint32_t mine = 0u;
#pragma unroll
    for (uint32_t r = 0; r < ROWS; ++r) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, (covered >> r) & 1u);
        if (lane == r) mine = m;
    }
    if (lane < nrows && mine != 0u) {
        const uint32_t lo = x0 + (uint32_t)__ffs((int)mine) - 1u, hic = W - 1u - (x0 + 31u - (uint32_t)__clz((int)mine));
        uint32_t *sp = row_spans + 2u * lane;
        if (lo < sp[0]) atomicMin(sp, lo);
        if (hic < sp[1]) atomicMin(sp + 1, hic);
    }
}

// Sparse delivery by the SMs: one warp per (row, frame) stores the row's covered span -- widened to 64-pixel columns, so that no cache
// line is shared with the host threads that write the background -- of the three colour planes and of the depth plane into the
// caller's mapped host buffers, 16 bytes per lane and step.  No per-rectangle copy-engine cost (~3 us per copy) and spans instead of
// boxes.  W % 16 == 0 and 16-byte aligned buffers (the host checks).
__global__ void __launch_bounds__(256) k_deliver(const uint32_t *__restrict__ spans, const uint8_t *__restrict__ rgb, const float *__restrict__ depth,
                                                 uint8_t *__restrict__ rgb_out, float *__restrict__ depth_out, uint32_t W, uint32_t rows, uint32_t n_frames) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= rows * n_frames) return;
    const uint32_t i = warp / rows, y = warp - i * rows;
    const uint32_t lo = spans[2u * warp], hic = spans[2u * warp + 1u];
    if (lo == 0xFFFFFFFFu) return;
    const uint32_t xa = lo & ~63u, xb = min(W, ((W - 1u - hic) | 63u) + 1u);
    const size_t P = (size_t)W * rows, row = (size_t)y * W + xa;
    const uint32_t n16 = (xb - xa) / 16u;
    if (rgb_out != nullptr) {
#pragma unroll
        for (uint32_t c = 0; c < 3u; ++c) {
            const uint4 *s = reinterpret_cast<const uint4 *>(rgb + ((size_t)i * 3u + c) * P + row);
            uint4 *d = reinterpret_cast<uint4 *>(rgb_out + ((size_t)i * 3u + c) * P + row);
            for (uint32_t k = lane; k < n16; k += 32u) d[k] = __ldcs(s + k);
        }
    }
    if (depth_out != nullptr) {
        const uint4 *s = reinterpret_cast<const uint4 *>(depth + (size_t)i * P + row);
        uint4 *d = reinterpret_cast<uint4 *>(depth_out + (size_t)i * P + row);
        for (uint32_t k = lane; k < n16 * 4u; k += 32u) d[k] = __ldcs(s + k);
    }
}

// Grid: x = 32-pixel tile columns, y = 16-row tile rows of the band, z = frame of the batch -- no thread divides to find its
// pixel.  A CTA owns one 32 x 16 tile (= one tile flag), each of its four warps four rows of it: lane = column.  (One warp per
// whole tile measured badly: a covered tile is 16 rows x ~330 instructions of serial work for one warp while its CTA's
// siblings over background have long exited and their slots idle -- a 640x480 frame's shade pass 15.6 -> 28.5 us, the 1080p
// spin batch only 4 % better than the row-segment kernel it replaced.)  Untouched tile: constant stores, done.  Touched tile:
// phase 1 issues the warp's four key loads (256 contiguous bytes per row) and keeps one coverage bit per row; phase 2 walks
// the rows, a covered pixel is shaded from the prepared record / gathered attributes, the results go to the warp's slice of
// shared memory and leave in two 16-byte store instructions (WIDE: W % 16 == 0 and 16-byte aligned planes; otherwise scalar
// stores from the loop, any W).  Planar output, CImg layout (CImg.h:11715-11721).
constexpr uint32_t SHADE_ROWS = SHADE_TILE_H / SHADE_WARPS; // rows per warp

// Output addresses of a warp's rows, computed from the block indices alone.  The touched path calls this AFTER its row loop
// with freshly (opaquely) re-read indices, so that no output pointer stays live across the shading code.
struct TileOut { uint8_t *rgb; float *depth; uint32_t nrows, x0; };
__device__ __forceinline__ TileOut tile_out(const View &vw, uint8_t *rgb, float *depth, uint32_t bx, uint32_t by, uint32_t bz, uint32_t warp) {
    TileOut t;
    t.x0 = bx * SHADE_TILE_W;
    const uint32_t r0 = by * SHADE_TILE_H + warp * SHADE_ROWS, rows = vw.y1 - vw.y0;
    t.nrows = r0 < rows ? min(SHADE_ROWS, rows - r0) : 0u;
    const size_t first = (size_t)r0 * vw.W + t.x0, P = vw.out_plane;
    t.rgb = rgb + (size_t)bz * 3 * P * vw.out_frame_stride + first;
    t.depth = depth ? depth + (size_t)bz * P * vw.out_frame_stride + first : nullptr;
    return t;
}

// A warp's four rows leave in 16-byte stores: colour 3 planes x 4 rows x 32 B = one store by 24 lanes, depth 4 rows x 128 B =
// one store by all lanes.  src_* = the warp's shared-memory slices ([3][4][32] bytes, [4][32] floats), or nullptr for the
// cleared frame (0 / 1.0f).
__device__ __forceinline__ void store_rows_wide(const TileOut &t, const View &vw, uint32_t lane, const uint8_t *src_rgb, const float *src_d) {
    const uint32_t W = vw.W;
    const uint32_t plane = lane >> 3, cr = (lane >> 1) & 3u, cx = (lane & 1u) * 16u;
    if (plane < 3u && cr < t.nrows && t.x0 + cx < W)
        *reinterpret_cast<uint4 *>(t.rgb + (size_t)plane * vw.out_plane + (size_t)cr * W + cx) = src_rgb ? reinterpret_cast<const uint4 *>(src_rgb)[lane] : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t dr = lane >> 3, dx = (lane & 7u) * 4u;
    if (t.depth && dr < t.nrows && t.x0 + dx < W)
        *reinterpret_cast<float4 *>(t.depth + (size_t)dr * W + dx) = src_d ? reinterpret_cast<const float4 *>(src_d)[lane] : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
}

#ifndef RAST_SHADE_MIN_BLOCKS
#define RAST_SHADE_MIN_BLOCKS 0
#endif
template <bool WIDE, bool PRE_NORMALS, bool FLAT, bool PREP>
#if RAST_SHADE_MIN_BLOCKS > 0
__global__ void __launch_bounds__(SHADE_WARPS * 32, RAST_SHADE_MIN_BLOCKS) k_resolve_shade(
#else
__global__ void __launch_bounds__(SHADE_WARPS * 32) k_resolve_shade(
#endif
    Scene sc, View vw, Batch bt, const __grid_constant__ LightTable lt, const LightDev *__restrict__ lights, uint8_t *__restrict__ rgb, float *__restrict__ depth, uint32_t keep_frame) {
    __shared__ __align__(16) uint8_t s_rgb[SHADE_WARPS][3][SHADE_ROWS * SHADE_TILE_W];
    __shared__ __align__(16) float s_depth[SHADE_WARPS][SHADE_ROWS * SHADE_TILE_W];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t x0 = blockIdx.x * SHADE_TILE_W, ty = blockIdx.y, f = blockIdx.z;
    const uint32_t W = vw.W, rows = vw.y1 - vw.y0, r0 = ty * SHADE_TILE_H + warp * SHADE_ROWS; // this warp's first row (band-relative)
    const uint32_t nrows = r0 < rows ? min(SHADE_ROWS, rows - r0) : 0u; // (a warp below the band's last row idles through: the CTA meets at a barrier)
    const uint32_t x = x0 + lane;
    const bool in_x = x < W;
    unsigned long long *vis = bt.vis + (size_t)f * vw.band_pixels + (size_t)r0 * W + x; // this lane's column of keys
    uint8_t *flag = bt.tile_flags + ((size_t)f * flag_tiles_y(vw) + ty) * flag_tiles_x(vw) + blockIdx.x;
    const bool reset = f != keep_frame; // hand the keys back as VIS_EMPTY (the next batch then needs no clear pass)

    bool touched = *flag != 0;
    uint32_t covered = 0; // bit r: this lane's pixel of row r has a winning triangle
    if (touched) {
        // phase 1: the key loads of the warp's rows in flight together.  The low word is the triangle index, 0xFFFFFFFF only in
        // VIS_EMPTY (rast_upload_mesh caps the triangle count below it).
#pragma unroll
        for (uint32_t r = 0; r < SHADE_ROWS; ++r)
            if (r < nrows && in_x) covered |= (*reinterpret_cast<const uint32_t *>(vis + (size_t)r * W) != INVALID_TRI ? 1u : 0u) << r;
        touched = __any_sync(0xFFFFFFFFu, covered != 0u);
    }
    if (reset) { // the flag goes back to 0 once every warp of the tile has read it
        __syncthreads();
        if (threadIdx.x == 0u) *flag = 0;
    }

    if (!touched) { // nothing drawn in these rows: the cleared frame (renderer.cpp:85-86)
        const TileOut t = tile_out(vw, rgb, depth, blockIdx.x, ty, f, warp);
        if (WIDE) {
            store_rows_wide(t, vw, lane, nullptr, nullptr);
        } else if (in_x) {
            const size_t P = vw.out_plane;
            for (uint32_t r = 0; r < nrows; ++r) {
                const size_t o = (size_t)r * W + lane;
                t.rgb[o] = 0; t.rgb[o + P] = 0; t.rgb[o + 2 * P] = 0;
                if (t.depth) t.depth[o] = 1.0f;
            }
        }
        return;
    }

    // Covered span of each of the warp's rows (sparse device-to-host delivery): one ballot per row, lane r keeps row r's and moves the
    // row's bounds only where it would (the plain read may be stale -- then the atomic is merely redundant).
    if (bt.spans != nullptr) note_row_spans<SHADE_ROWS>(bt.spans + ((size_t)f * rows + r0) * 2u, covered, nrows, lane, x0, W);

    // phase 2.  The per-frame base pointers are made opaque so that a gather is "base + index * 16" (one
    // IMAD.WIDE) instead of a 64-bit add of the frame offset to every index followed by the address computation.
    unsigned long long rv_base = (unsigned long long)(bt.rv + (size_t)f * sc.V);
    unsigned long long cn_base = (unsigned long long)(PRE_NORMALS ? bt.cn + (size_t)f * sc.Nn : nullptr);
    asm volatile("" : "+l"(rv_base), "+l"(cn_base));
    const float4 *rv = reinterpret_cast<const float4 *>(rv_base);
    const float4 *cn = reinterpret_cast<const float4 *>(cn_base);
    const FrameParams *fp = bt.frames + f;
#if RAST_SHADE_PREP
    unsigned long long prep_base = (unsigned long long)(PREP ? bt.prep + (size_t)f * sc.T * PREP_QUADS : nullptr);
    asm volatile("" : "+l"(prep_base));
    const float4 *prep = reinterpret_cast<const float4 *>(prep_base);
#endif
    const bool cw = __ldg(&fp->wind_clockwise) != 0u;
#if RAST_SHADE_PREP
    PrepRec rec;
#endif
    uint8_t *sr = &s_rgb[warp][0][lane];
    float *sd = &s_depth[warp][lane];
    TileOut tn{}; // non-WIDE: scalar stores from the loop
    if (!WIDE) tn = tile_out(vw, rgb, depth, blockIdx.x, ty, f, warp);
#pragma unroll 1
    for (uint32_t r = 0; r < nrows; ++r) {
        Shaded px;
        px.r = px.g = px.b = 0u; px.depth = 1.0f; // renderer.cpp:85-86
        if ((covered >> r) & 1u) {
            unsigned long long *key = vis + (size_t)r * W;
            const uint32_t tri = *reinterpret_cast<const uint32_t *>(key); // L1 / L2 hit: phase 1 fetched the line
#if RAST_SHADE_PREP
            if (PREP) {
                load_prep_record(rec, tri, sc, prep);
                px = shade_prepared(rec, x, vw.y0 + r0 + r, cw, lt, lights);
            } else
#endif
            px = shade_pixel<PRE_NORMALS, FLAT>(tri, x, vw.y0 + r0 + r, sc, rv, cn, FLAT ? fp->modelview : fp->normal_m, cw, lt, lights);
            if (reset) *key = VIS_EMPTY;
        }
        if (WIDE) {
            sr[r * SHADE_TILE_W] = (uint8_t)px.r;
            sr[r * SHADE_TILE_W + SHADE_ROWS * SHADE_TILE_W] = (uint8_t)px.g;
            sr[r * SHADE_TILE_W + 2 * SHADE_ROWS * SHADE_TILE_W] = (uint8_t)px.b;
            sd[r * SHADE_TILE_W] = px.depth;
        } else if (in_x) {
            const size_t o = (size_t)r * W + lane, P = vw.out_plane;
            tn.rgb[o] = (uint8_t)px.r; tn.rgb[o + P] = (uint8_t)px.g; tn.rgb[o + 2 * P] = (uint8_t)px.b;
            if (tn.depth) tn.depth[o] = px.depth;
        }
    }
    if (WIDE) {
        __syncwarp();
        uint32_t bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z, wp = threadIdx.x >> 5;
        asm volatile("" : "+r"(bx), "+r"(by), "+r"(bz), "+r"(wp)); // opaque: recomputed here, not carried through the loop
        const TileOut t = tile_out(vw, rgb, depth, bx, by, bz, wp);
        store_rows_wide(t, vw, threadIdx.x & 31u, &s_rgb[wp][0][0], &s_depth[wp][0]);
    }
}

// ---- the same pass with one warp per whole tile (big batches) -------------------------------------------------------------
// k_resolve_shade above gives every warp 4 rows: the most parallel slices, right for a single small frame (a 640x480 frame is
// 600 tiles).  A big batch (32 frames of 1080p = 130 k tiles) has parallelism to spare and is better served by amortising a
// tile's latency chain (flag -> keys -> records) over 16 rows: one warp per 32 x 16 tile.  Warps must then not share a CTA's
// fate: a covered tile is ~5 k instructions, a background one ~40, and a CTA holds its slots until its slowest warp is done.
// RAST_SHADE_CTA_WARPS = warps per CTA (side by side in x); RAST_SHADE_PERSIST = 1: a persistent grid whose warps fetch tiles
// from a counter (RAST_SHADE_GRAB consecutive tiles per fetch) instead of one CTA per RAST_SHADE_CTA_WARPS tiles.
#ifndef RAST_SHADE_CTA_WARPS
#define RAST_SHADE_CTA_WARPS 4
#endif
#ifndef RAST_SHADE_PERSIST
#define RAST_SHADE_PERSIST 0
#endif
#ifndef RAST_SHADE_GRAB
#define RAST_SHADE_GRAB 4
#endif
constexpr uint32_t SHADE_WT_WARPS = RAST_SHADE_CTA_WARPS;

struct WarpTile { uint32_t f, ty, tx; };
__device__ __forceinline__ WarpTile decode_tile(const View &vw, uint32_t t) { // t = (f * tiles_y + ty) * tiles_x + tx
    const uint32_t tx_n = flag_tiles_x(vw), ty_n = flag_tiles_y(vw);
    WarpTile w;
    w.tx = t % tx_n;
    const uint32_t q = t / tx_n;
    w.ty = q % ty_n;
    w.f = q / ty_n;
    return w;
}

__device__ __forceinline__ void store_tile_wide(const View &vw, uint8_t *rgb, float *depth, const WarpTile &w, uint32_t lane, const uint8_t *src_rgb, const float *src_d) {
    const uint32_t W = vw.W, x0 = w.tx * SHADE_TILE_W, r0 = w.ty * SHADE_TILE_H, nrows = min(SHADE_TILE_H, vw.y1 - vw.y0 - r0);
    const size_t first = (size_t)r0 * W + x0, P = vw.out_plane;
    uint8_t *o_rgb = rgb + (size_t)w.f * 3 * P * vw.out_frame_stride + first;
    const uint32_t cr = lane >> 1, cx = (lane & 1u) * 16u; // colour planes: 16 rows x 32 B = one 16-byte store per lane and plane
    if (cr < nrows && x0 + cx < W) {
        uint8_t *p = o_rgb + (size_t)cr * W + cx;
        constexpr uint32_t PL = SHADE_TILE_W * SHADE_TILE_H;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(p) = src_rgb ? reinterpret_cast<const uint4 *>(src_rgb)[lane] : z;
        *reinterpret_cast<uint4 *>(p + P) = src_rgb ? reinterpret_cast<const uint4 *>(src_rgb + PL)[lane] : z;
        *reinterpret_cast<uint4 *>(p + 2 * P) = src_rgb ? reinterpret_cast<const uint4 *>(src_rgb + 2 * PL)[lane] : z;
    }
    if (depth) { // 16 rows x 128 B = four 16-byte stores per lane
        float *o_d = depth + (size_t)w.f * P * vw.out_frame_stride + first;
        const uint32_t dx = (lane & 7u) * 4u;
        const float4 one = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
#pragma unroll
        for (uint32_t j = 0; j < 4u; ++j) {
            const uint32_t dr = j * 4u + (lane >> 3);
            if (dr < nrows && x0 + dx < W) *reinterpret_cast<float4 *>(o_d + (size_t)dr * W + dx) = src_d ? reinterpret_cast<const float4 *>(src_d)[j * 32u + lane] : one;
        }
    }
}

template <bool WIDE, bool PRE_NORMALS, bool FLAT, bool PREP>
__device__ __forceinline__ void shade_warp_tile(uint32_t t, uint32_t lane, uint8_t *s_rgb, float *s_depth, const Scene &sc, const View &vw, const Batch &bt,
                                                const LightTable &lt, const LightDev *__restrict__ lights, uint8_t *__restrict__ rgb, float *__restrict__ depth, uint32_t keep_frame) {
    const WarpTile w = decode_tile(vw, t);
    const uint32_t f = w.f;
    const uint32_t W = vw.W, rows = vw.y1 - vw.y0, x0 = w.tx * SHADE_TILE_W, r0 = w.ty * SHADE_TILE_H;
    const uint32_t nrows = min(SHADE_TILE_H, rows - r0);
    const uint32_t x = x0 + lane;
    const bool in_x = x < W;
    unsigned long long *vis = bt.vis + (size_t)f * vw.band_pixels + (size_t)r0 * W + x; // this lane's column of keys
    uint8_t *flag = bt.tile_flags + t;
    const bool reset = f != keep_frame; // hand the keys back as VIS_EMPTY (the next batch then needs no clear pass)

    bool touched = *flag != 0;
    uint32_t covered = 0; // bit r: this lane's pixel of row r has a winning triangle
    if (touched) {
        // phase 1: all key loads in flight at once.  The low word is the triangle index, 0xFFFFFFFF only in VIS_EMPTY
        // (rast_upload_mesh caps the triangle count below it).
#pragma unroll
        for (uint32_t r = 0; r < SHADE_TILE_H; ++r)
            if (r < nrows && in_x) covered |= (*reinterpret_cast<const uint32_t *>(vis + (size_t)r * W) != INVALID_TRI ? 1u : 0u) << r;
        touched = __any_sync(0xFFFFFFFFu, covered != 0u);
        if (reset && lane == 0u) *flag = 0;
    }

    if (!touched) { // nothing drawn here: the cleared frame (renderer.cpp:85-86)
        if (WIDE) {
            store_tile_wide(vw, rgb, depth, w, lane, nullptr, nullptr);
        } else if (in_x) {
            const size_t P = vw.out_plane, first = (size_t)r0 * W + x0;
            uint8_t *o_rgb = rgb + (size_t)f * 3 * P * vw.out_frame_stride + first;
            float *o_d = depth ? depth + (size_t)f * P * vw.out_frame_stride + first : nullptr;
            for (uint32_t r = 0; r < nrows; ++r) {
                const size_t o = (size_t)r * W + lane;
                o_rgb[o] = 0; o_rgb[o + P] = 0; o_rgb[o + 2 * P] = 0;
                if (o_d) o_d[o] = 1.0f;
            }
        }
        return;
    }

    if (bt.spans != nullptr) note_row_spans<SHADE_TILE_H>(bt.spans + ((size_t)f * rows + r0) * 2u, covered, nrows, lane, x0, W); // as in k_resolve_shade

    unsigned long long rv_base = (unsigned long long)(bt.rv + (size_t)f * sc.V);
    unsigned long long cn_base = (unsigned long long)(PRE_NORMALS ? bt.cn + (size_t)f * sc.Nn : nullptr);
    asm volatile("" : "+l"(rv_base), "+l"(cn_base));
    const float4 *rv = reinterpret_cast<const float4 *>(rv_base);
    const float4 *cn = reinterpret_cast<const float4 *>(cn_base);
    const FrameParams *fp = bt.frames + f;
#if RAST_SHADE_PREP
    unsigned long long prep_base = (unsigned long long)(PREP ? bt.prep + (size_t)f * sc.T * PREP_QUADS : nullptr);
    asm volatile("" : "+l"(prep_base));
    const float4 *prep = reinterpret_cast<const float4 *>(prep_base);
#endif
    const bool cw = __ldg(&fp->wind_clockwise) != 0u;
#if RAST_SHADE_PREP
    PrepRec rec;
#endif
    uint8_t *sr = s_rgb + lane;
    float *sd = s_depth + lane;
    uint8_t *n_rgb = nullptr; // non-WIDE: scalar stores from the loop
    float *n_d = nullptr;
    if (!WIDE) {
        const size_t P = vw.out_plane, first = (size_t)r0 * W + x0;
        n_rgb = rgb + (size_t)f * 3 * P * vw.out_frame_stride + first;
        n_d = depth ? depth + (size_t)f * P * vw.out_frame_stride + first : nullptr;
    }
#pragma unroll 1
    for (uint32_t r = 0; r < nrows; ++r) {
        Shaded px;
        px.r = px.g = px.b = 0u; px.depth = 1.0f; // renderer.cpp:85-86
        if ((covered >> r) & 1u) {
            unsigned long long *key = vis + (size_t)r * W;
            const uint32_t tri = *reinterpret_cast<const uint32_t *>(key); // L1 / L2 hit: phase 1 fetched the line
#if RAST_SHADE_PREP
            if (PREP) {
                load_prep_record(rec, tri, sc, prep);
                px = shade_prepared(rec, x, vw.y0 + r0 + r, cw, lt, lights);
            } else
#endif
            px = shade_pixel<PRE_NORMALS, FLAT>(tri, x, vw.y0 + r0 + r, sc, rv, cn, FLAT ? fp->modelview : fp->normal_m, cw, lt, lights);
            if (reset) *key = VIS_EMPTY;
        }
        if (WIDE) {
            sr[r * SHADE_TILE_W] = (uint8_t)px.r;
            sr[r * SHADE_TILE_W + SHADE_TILE_W * SHADE_TILE_H] = (uint8_t)px.g;
            sr[r * SHADE_TILE_W + 2 * SHADE_TILE_W * SHADE_TILE_H] = (uint8_t)px.b;
            sd[r * SHADE_TILE_W] = px.depth;
        } else if (in_x) {
            const size_t o = (size_t)r * W + lane, P = vw.out_plane;
            n_rgb[o] = (uint8_t)px.r; n_rgb[o + P] = (uint8_t)px.g; n_rgb[o + 2 * P] = (uint8_t)px.b;
            if (n_d) n_d[o] = px.depth;
        }
    }
    if (WIDE) {
        __syncwarp();
        uint32_t t2 = t;
        asm volatile("" : "+r"(t2)); // opaque: the output addresses are recomputed here, not carried through the loop
        store_tile_wide(vw, rgb, depth, decode_tile(vw, t2), lane, s_rgb, s_depth);
        __syncwarp();
    }
}

// Measured on a B200 (120-frame 1080p calls, shade pass per call, one session): 1 warp per CTA 1.71 ms, 2: 1.61, 4: 1.57, 4 with
// __launch_bounds__(128, 6): 1.49 -- the hint lets ptxas take 72 registers instead of 64 (28 instead of 32 resident warps, but fewer
// serialised dependent chains per pixel); (128, 8) is the 64-register code again.
#ifndef RAST_SHADE_WT_MIN_BLOCKS
#define RAST_SHADE_WT_MIN_BLOCKS 6
#endif
template <bool WIDE, bool PRE_NORMALS, bool FLAT, bool PREP>
#if RAST_SHADE_WT_MIN_BLOCKS > 0
__global__ void __launch_bounds__(SHADE_WT_WARPS * 32, RAST_SHADE_WT_MIN_BLOCKS) k_resolve_shade_wt(
#else
__global__ void __launch_bounds__(SHADE_WT_WARPS * 32) k_resolve_shade_wt(
#endif
    Scene sc, View vw, Batch bt, const __grid_constant__ LightTable lt, const LightDev *__restrict__ lights, uint8_t *__restrict__ rgb, float *__restrict__ depth,
    uint32_t keep_frame, uint32_t n_tiles, unsigned int *cursor) {
    __shared__ __align__(16) uint8_t s_rgb[SHADE_WT_WARPS][3 * SHADE_TILE_W * SHADE_TILE_H];
    __shared__ __align__(16) float s_depth[SHADE_WT_WARPS][SHADE_TILE_W * SHADE_TILE_H];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#if RAST_SHADE_PERSIST
    for (;;) {
        uint32_t base = 0;
        if (lane == 0u) base = atomicAdd(cursor, (unsigned int)RAST_SHADE_GRAB);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n_tiles) break;
        const uint32_t end = min(base + (uint32_t)RAST_SHADE_GRAB, n_tiles);
        for (uint32_t t = base; t < end; ++t) shade_warp_tile<WIDE, PRE_NORMALS, FLAT, PREP>(t, lane, s_rgb[warp], s_depth[warp], sc, vw, bt, lt, lights, rgb, depth, keep_frame);
    }
#else
    const uint32_t t = blockIdx.x * SHADE_WT_WARPS + warp;
    if (t < n_tiles) shade_warp_tile<WIDE, PRE_NORMALS, FLAT, PREP>(t, lane, s_rgb[warp], s_depth[warp], sc, vw, bt, lt, lights, rgb, depth, keep_frame);
#endif
}

// ---- auxiliary kernels ----------------------------------------------------------------------
// Self-test of exact::div3 against __fdiv_rn: pseudo-random operand pairs (xorshift), exponents spread over and
// beyond the guarded range, mantissas biased towards the patterns that stress rounding (all ones, single bits,
// near powers of two).  Counts quotients whose bits differ.
__global__ void k_selftest_division(unsigned long long samples_per_thread, unsigned long long seed, unsigned long long *mismatches) {
    unsigned long long x = seed ^ (0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1ull));
    unsigned long long bad = 0;
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    auto make = [&](unsigned long long r) {
        uint32_t mant = (uint32_t)(r >> 8) & 0x7FFFFFu;
        const uint32_t style = (uint32_t)(r >> 40) & 7u;
        if (style == 0u) mant = 0x7FFFFFu ^ ((uint32_t)(r >> 44) & 0xFu);      // all ones, last bits flipped
        else if (style == 1u) mant = 1u << ((uint32_t)(r >> 44) % 23u);          // a single bit
        else if (style == 2u) mant = (uint32_t)(r >> 44) & 0xFu;                  // just above a power of two
        const uint32_t expo = 127u - 60u + (uint32_t)((r >> 48) % 121ull);       // 2^-60 .. 2^60
        return __uint_as_float(((uint32_t)(r >> 63) << 31) | (expo << 23) | mant);
    };
    for (unsigned long long i = 0; i < samples_per_thread; ++i) {
        const float b = make(next()), a0 = make(next()), a1 = make(next());
        const float a2 = (i & 63ull) == 0ull ? 0.0f : make(next()); // zeros take the fallback path
        float q0, q1, q2;
        exact::div3(a0, a1, a2, b, exact::div_reciprocal(b), exact::div_in_range(b), q0, q1, q2);
        bad += __float_as_uint(q0) != __float_as_uint(__fdiv_rn(a0, b));
        bad += __float_as_uint(q1) != __float_as_uint(__fdiv_rn(a1, b));
        bad += __float_as_uint(q2) != __float_as_uint(__fdiv_rn(a2, b));
    }
    if (bad) atomicAdd(mismatches, bad);
}

// rast_upload_mesh on the device: the caller's Triangle array (10 x int32 per triangle, headers/face.h:6-13) is copied
// to the GPU as it is and rearranged here -- vertex indices into three SoA arrays, the rest into the 48-byte record of
// the shade pass, absent (-1) normals / uvs pointed at the sentinel entry.  Out-of-range indices (undefined behaviour in
// the reference: vector operator[], drawing.cpp:167-173) are reported through `first_error`: the smallest key
// (triangle * 16 + corner * 4 + kind) wins, i.e. the error a sequential check in triangle order would hit first.
enum { UPLOAD_ERR_VERTEX = 0, UPLOAD_ERR_NORMAL = 1, UPLOAD_ERR_UV = 2 };
__global__ void __launch_bounds__(256) k_build_tri_records(const int *__restrict__ tris, uint64_t n_tris, uint32_t n_positions, uint32_t n_normals, uint32_t n_uvs,
                                                            int *__restrict__ vidx, int4 *__restrict__ rec, unsigned long long *first_error) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const int2 *src = reinterpret_cast<const int2 *>(tris + 10 * t); // 40-byte rows are 8-byte aligned
    int f[10];
#pragma unroll
    for (int k = 0; k < 5; ++k) { const int2 v = __ldg(src + k); f[2 * k] = v.x; f[2 * k + 1] = v.y; }
    int n[3], uvi[3];
    unsigned long long err = ~0ull;
#pragma unroll
    for (int k = 2; k >= 0; --k) { // descending, so that the smallest key is kept
        if (f[6 + k] >= 0 && (uint32_t)f[6 + k] >= n_uvs) err = t * 16ull + k * 4 + UPLOAD_ERR_UV;
        if (f[3 + k] >= 0 && (uint32_t)f[3 + k] >= n_normals) err = t * 16ull + k * 4 + UPLOAD_ERR_NORMAL;
        if (f[k] < 0 || (uint32_t)f[k] >= n_positions) err = t * 16ull + k * 4 + UPLOAD_ERR_VERTEX;
        n[k] = f[3 + k] >= 0 ? f[3 + k] : (int)n_normals;
        uvi[k] = f[6 + k] >= 0 ? f[6 + k] : (int)n_uvs;
    }
    if (err != ~0ull) atomicMin(first_error, err);
    vidx[t] = f[0]; vidx[n_tris + t] = f[1]; vidx[2 * n_tris + t] = f[2];
    rec[3 * t] = make_int4(f[0], f[1], f[2], n[0]);
    rec[3 * t + 1] = make_int4(n[1], n[2], uvi[0], uvi[1]);
    rec[3 * t + 2] = make_int4(uvi[2], f[9], f[9], 0); // .z keeps the caller's material index, .y is resolved at draw time
}

// tri_rec[3t+2].z holds the caller's material index; .y becomes the index the shade pass uses:
// materials[face.material] (drawing.cpp:173), with -1 / out of range mapped to the sentinel.
__global__ void k_resolve_materials(int4 *tri_rec, uint64_t n_tris, uint32_t n_materials) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    int4 r = tri_rec[3 * t + 2];
    r.y = (r.z < 0 || (uint32_t)r.z >= n_materials) ? (int)n_materials : r.z;
    tri_rec[3 * t + 2] = r;
}

// xyz normals -> (x, y, z, 0) at 16-byte stride for the shade pass (once per upload; entry n is the zero sentinel)
__global__ void k_pad_normals(const float *__restrict__ nrm, float4 *__restrict__ nrm4, uint32_t n_with_sentinel) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_with_sentinel) nrm4[i] = make_float4(nrm[3 * (size_t)i], nrm[3 * (size_t)i + 1], nrm[3 * (size_t)i + 2], 0.f);
}

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
