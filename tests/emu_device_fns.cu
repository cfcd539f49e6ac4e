// emu_device_fns.cu -- TEST ONLY.  Runs the per-thread functions of rasteriser_b200/csrc/kernels.cuh ON THE HOST.
//
// The development container has no GPU.  The arithmetic of the frame path lives in RAST_HD (__host__ __device__) functions
// -- raster_vertex, signed_area_2d, bounding_box, tri_setup, edges / candidate / fragment, stage_item / raster_item (the warp
// rasteriser's inner loop, one lane at a time), shade_pixel, sample_texture, and the variants' rast_tight_bbox, prepare_triangle / shade_pixel_prep -- so this file, compiled by nvcc for the host,
// drives those very functions serially in the order the kernels launch them and tests/test_emu_device_fns.py compares the
// result with the oracle bit for bit.  What it does NOT cover: the __global__ wrappers (grids, staging through shared
// memory, votes, atomics, streams) and the device flavour of exact:: (the _rn intrinsics and the shared-reciprocal division,
// which rast_selftest_division checks on the GPU).  It is a checker for kernel-logic changes made without a GPU, never a
// product path: nothing under rasteriser_b200/ or include/ builds, loads or calls it.
//
// Build (tests/test_emu_device_fns.py): nvcc -std=c++17 -O2 -Xcompiler -fPIC,-ffp-contract=off -shared -o libemu.so emu_device_fns.cu
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

// ---- a host warp in lockstep (flag EMU_WARP): 32 threads, one per lane; votes and warp reductions are real -------------------
// Every lane of raster_item reaches the same collectives in the same order (they sit in warp-uniform control flow), so a
// barrier per collective is a faithful model of __any_sync / __reduce_max_sync.  Without EMU_WARP the lanes run one after the
// other and a collective degenerates to the lane's own value (still a valid, merely different, execution).
namespace emu_warp {
struct Ctx {
    std::atomic<int> arrived{0}, sense{0};
    uint32_t vals[32];
};
thread_local Ctx *g_ctx = nullptr;
thread_local uint32_t g_lane = 0;
inline void barrier(Ctx *c, int parties) {
    const int s = c->sense.load();
    if (c->arrived.fetch_add(1) == parties - 1) { c->arrived.store(0); c->sense.store(s ^ 1); }
    else while (c->sense.load() == s) std::this_thread::yield();
}
inline uint32_t warp_max(uint32_t v) {
    Ctx *c = g_ctx;
    if (!c) return v;
    c->vals[g_lane] = v;
    barrier(c, 32);
    uint32_t m = 0;
    for (int i = 0; i < 32; ++i) m = c->vals[i] > m ? c->vals[i] : m;
    barrier(c, 32); // nobody overwrites vals before everybody has read them
    return m;
}
inline bool any(bool p) { return warp_max(p ? 1u : 0u) != 0u; }
} // namespace emu_warp
#define RAST_HOST_ANY(p) emu_warp::any(p)
#define RAST_HOST_WARP_MAX_U32(v) emu_warp::warp_max(v)

#ifndef RAST_TIGHT_TINY
#define RAST_TIGHT_TINY 1 // the host driver below calls rast_tight_bbox only when asked to (flag bit 0)
#endif
#ifndef RAST_SHADE_PREP
#define RAST_SHADE_PREP 1 // prepare_triangle / shade_pixel_prep are driven only when asked to (flag bit 5)
#endif
#include "../rasteriser_b200/csrc/kernels.cuh"

struct EmuMaterial { // = rast_material (include/rast.h): planar normalised texels
    float kd[3];
    int32_t has_texture, tex_w, tex_h;
    const float *texels;
};

enum { EMU_TIGHT = 1, EMU_PRE_NORMALS = 2, EMU_EARLY_Z = 4, EMU_ALL_CHUNKS = 8, EMU_FLAT_FACE = 16, EMU_PREP = 32, EMU_WARP = 64, EMU_TILES = 128 };

// One frame.  lights: n x 10 floats (rast_light: direction, intensity, colour, trans_dir -- trans_dir already computed by
// rast_transform_lights).  Outputs: rgb planar [3][rows][W], depth [rows][W], tri_ids [rows][W] for the band [y0, y1).
extern "C" int emu_draw(const float *pos, uint32_t V, const float *nrm, uint32_t Nn, const float *uv, uint32_t Nuv, const int32_t *tris, uint64_t T,
                        const EmuMaterial *mats, uint32_t M, const float *lights, uint32_t L, const float *camera, const float *normal_m,
                        const float *modelview, int wind_clockwise, uint32_t W, uint32_t H, uint32_t y0, uint32_t y1, uint32_t tiny_max_pixels, int flags,
                        uint8_t *rgb, float *depth, uint32_t *tri_ids) {
    using namespace rk;
    View vw;
    vw.W = W; vw.H = H; vw.y0 = y0; vw.y1 = y1; vw.band_pixels = W * (y1 - y0); vw.out_plane = vw.band_pixels;
    const size_t P = vw.band_pixels;

    // ---- scene arrays as rast_upload_mesh / k_build_tri_records / rast_upload_materials lay them out ----
    std::vector<float> nrm_s(nrm, nrm + 3 * (size_t)Nn);
    nrm_s.insert(nrm_s.end(), {0.f, 0.f, 0.f}); // sentinel normal
    std::vector<float2> uv_s((size_t)Nuv + 1);
    for (uint32_t i = 0; i < Nuv; ++i) uv_s[i] = make_float2(uv[2 * i], uv[2 * i + 1]);
    uv_s[Nuv] = make_float2(0.f, 0.f);
    std::vector<int4> rec(3 * (size_t)T);
    for (uint64_t t = 0; t < T; ++t) {
        const int32_t *f = tris + 10 * t;
        int n[3], u[3];
        for (int k = 0; k < 3; ++k) {
            if (f[k] < 0 || (uint32_t)f[k] >= V) return -1;
            if (f[3 + k] >= 0 && (uint32_t)f[3 + k] >= Nn) return -1;
            if (f[6 + k] >= 0 && (uint32_t)f[6 + k] >= Nuv) return -1;
            n[k] = f[3 + k] >= 0 ? f[3 + k] : (int)Nn;
            u[k] = f[6 + k] >= 0 ? f[6 + k] : (int)Nuv;
        }
        const int mat = (f[9] < 0 || (uint32_t)f[9] >= M) ? (int)M : f[9]; // k_resolve_materials
        rec[3 * t] = make_int4(f[0], f[1], f[2], n[0]);
        rec[3 * t + 1] = make_int4(n[1], n[2], u[0], u[1]);
        rec[3 * t + 2] = make_int4(u[2], mat, f[9], 0);
    }
    std::vector<MaterialDev> md((size_t)M + 1);
    std::vector<float4> texels;
    for (uint32_t i = 0; i < M; ++i) {
        md[i].kd[0] = mats[i].kd[0]; md[i].kd[1] = mats[i].kd[1]; md[i].kd[2] = mats[i].kd[2];
        md[i].has_texture = mats[i].has_texture ? (1 | (mats[i].has_texture & 2)) : 0;
        md[i].tex_w = mats[i].tex_w; md[i].tex_h = mats[i].tex_h;
        md[i].texel_offset = (long long)texels.size();
        if (mats[i].has_texture) {
            const size_t n = (size_t)mats[i].tex_w * mats[i].tex_h;
            const float *t = mats[i].texels;
            for (size_t k = 0; k < n; ++k) texels.push_back(make_float4(t[k], t[n + k], t[2 * n + k], 0.f));
        }
    }
    md[M].kd[0] = md[M].kd[1] = md[M].kd[2] = 1.f;
    md[M].has_texture = 0; md[M].tex_w = md[M].tex_h = 0; md[M].texel_offset = 0;
    texels.push_back(make_float4(0.f, 0.f, 0.f, 0.f));

    std::vector<float4> nrm4((size_t)Nn + 1); // what k_pad_normals writes at upload
    for (size_t i = 0; i <= (size_t)Nn; ++i) nrm4[i] = make_float4(nrm_s[3 * i], nrm_s[3 * i + 1], nrm_s[3 * i + 2], 0.f);
    Scene sc{};
    sc.pos = pos; sc.nrm = nrm_s.data(); sc.nrm4 = nrm4.data(); sc.uv = uv_s.data();
    sc.tri_rec = rec.data(); sc.mats = md.data(); sc.texels = texels.data();
    sc.V = V; sc.Nn = Nn + 1; sc.Nuv = Nuv + 1; sc.M = M + 1; sc.T = T;

    // ---- lights as draw_frames_impl prepares them ----
    LightTable lt{};
    std::vector<LightDev> ld((size_t)L + 1);
    for (uint32_t l = 0; l < L; ++l) {
        const float *q = lights + 10 * (size_t)l;
        ld[l].ntx = -q[7]; ld[l].nty = -q[8]; ld[l].ntz = -q[9];
        ld[l].icr = q[3] * q[4]; ld[l].icg = q[3] * q[5]; ld[l].icb = q[3] * q[6];
        ld[l].pad0 = ld[l].pad1 = 0.f;
        if (l < PARAM_LIGHTS) { lt.a[l] = make_float4(ld[l].ntx, ld[l].nty, ld[l].ntz, ld[l].icr); lt.c[l] = make_float2(ld[l].icg, ld[l].icb); }
    }
    lt.n = L;

    // ---- k_vertex ----
    std::vector<float4> rv(V), cn((size_t)Nn + 1);
    for (uint32_t i = 0; i < V; ++i) rv[i] = raster_vertex(camera, pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2], W, H);
    for (uint32_t j = 0; j <= Nn; ++j) {
        const float4 t = exact::mat_vec(normal_m, nrm_s[3 * (size_t)j], nrm_s[3 * (size_t)j + 1], nrm_s[3 * (size_t)j + 2], 0.f);
        cn[j] = make_float4(t.x, t.y, t.z, 0.f);
    }

    // ---- k_setup (+ k_raster_chunks): cull, bbox, tiny bboxes walked by edges / candidate / fragment, the rest cut into
    //      CHUNK x CHUNK items that go through stage_item / raster_item lane by lane ----
    std::vector<unsigned long long> vis(P, VIS_EMPTY);
    StagedTris *stg = new StagedTris();
    const bool early_z = (flags & EMU_EARLY_Z) != 0;
    // EMU_WARP: 32 lane threads wait for the main thread to stage an item, rasterise it in lockstep, and report back
    emu_warp::Ctx warp_ctx, gate; // warp_ctx: collectives among the 32 lanes; gate: main thread + 32 lanes (33 parties)
    std::atomic<int> warp_slot{-1};
    std::vector<std::thread> lanes;
    if (flags & EMU_WARP) {
        for (uint32_t lane = 0; lane < 32u; ++lane)
            lanes.emplace_back([&, lane]() {
                emu_warp::g_ctx = &warp_ctx;
                emu_warp::g_lane = lane;
                for (;;) {
                    emu_warp::barrier(&gate, 33); // item staged (or stop)
                    const int slot = warp_slot.load();
                    if (slot < 0) return;
                    raster_item<false>(*stg, (uint32_t)slot, lane, vw, vis.data(), nullptr, early_z);
                    emu_warp::barrier(&gate, 33); // item done
                }
            });
    }
    std::vector<uint32_t> tile_list;
    for (uint64_t t = 0; t < T; ++t) {
        const float4 v0 = rv[tris[10 * t]], v1 = rv[tris[10 * t + 1]], v2 = rv[tris[10 * t + 2]];
        const float a2 = signed_area_2d(v0, v1, v2);
        if (!((a2 > 0.f) != (wind_clockwise != 0))) continue;
        const BBox bb = bounding_box(v0, v1, v2, vw);
        if (bb.empty) continue;
        const uint32_t w = bb.x1 - bb.x0 + 1u, h = bb.y1 - bb.y0 + 1u;
        const bool inline_raster = !(flags & EMU_ALL_CHUNKS) && (uint64_t)w * h <= tiny_max_pixels;
        if (inline_raster) {
            TriSetup s;
            tri_setup(s, v0, v1, v2);
            uint32_t wx0 = bb.x0, wy0 = bb.y0, wx1 = bb.x1, wy1 = bb.y1;
            if ((flags & EMU_TIGHT) && !rast_tight_bbox(v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, s.literal ? 0.f : s.area, &wx0, &wy0, &wx1, &wy1)) continue;
            for (uint32_t y = wy0; y <= wy1; ++y)
                for (uint32_t x = wx0; x <= wx1; ++x) test_and_commit(s, x, y, (uint32_t)t, vis.data(), vw);
        } else if (flags & EMU_TILES) {
            tile_list.push_back((uint32_t)t); // the screen-tile schedule: binned below
        } else {
            const uint32_t ncx = (w + CHUNK - 1) / CHUNK, ncy = (h + CHUNK - 1) / CHUNK;
            for (uint32_t cy = 0; cy < ncy; ++cy)
                for (uint32_t cx = 0; cx < ncx; ++cx) {
                    const uint32_t rx0 = bb.x0 + cx * CHUNK, ry0 = bb.y0 + cy * CHUNK;
                    const uint32_t rx1 = min(bb.x1, rx0 + CHUNK - 1u), ry1 = min(bb.y1, ry0 + CHUNK - 1u);
                    const uint32_t slot = (uint32_t)((t + cx + cy) & 31u); // any staging slot must do
                    stage_item(*stg, slot, (uint32_t)t, 0u, v0, v1, v2, rx0, ry0, rx1, ry1, rx0, ry0);
                    if (flags & EMU_WARP) {
                        warp_slot.store((int)slot);
                        emu_warp::barrier(&gate, 33);
                        emu_warp::barrier(&gate, 33);
                    } else {
                        for (uint32_t lane = 0; lane < 32u; ++lane) raster_item<false>(*stg, slot, lane, vw, vis.data(), nullptr, early_z);
                    }
                }
        }
    }
    if (flags & EMU_WARP) {
        warp_slot.store(-1);
        emu_warp::barrier(&gate, 33);
        for (std::thread &t : lanes) t.join();
    }
#if RAST_BLOCK_Z
    if (flags & EMU_TILES) {
        // ---- k_fill_tiles / k_raster_tiles, one tile after the other: the bin near to far (nearest-vertex depth key), the tile's keys in a
        //      private array, the farthest stored depth of each 16 x 8 block refreshed after every item, an item's eight blocks tested against
        //      them with the kernel's own block_behind (the vote of k_raster_tiles: lane b = block b) and the survivors handed to
        //      raster_item<tile schedule> lane by lane with that mask; finally the tile is merged into the visibility buffer ----
        const uint32_t tiles_x = (W + TILE - 1) / TILE, tiles_y = (y1 - y0 + TILE - 1) / TILE;
        std::vector<std::vector<std::pair<uint32_t, uint32_t>>> bins((size_t)tiles_x * tiles_y); // (nearest depth key, triangle)
        for (uint32_t t : tile_list) {
            const float4 v0 = rv[tris[10 * (size_t)t]], v1 = rv[tris[10 * (size_t)t + 1]], v2 = rv[tris[10 * (size_t)t + 2]];
            const BBox bb = bounding_box(v0, v1, v2, vw);
            const uint32_t zkey = depth_key(fminf(fminf(v0.z, v1.z), v2.z));
            for (uint32_t ty = (bb.y0 - y0) / TILE; ty <= (bb.y1 - y0) / TILE; ++ty)
                for (uint32_t tx = bb.x0 / TILE; tx <= bb.x1 / TILE; ++tx) bins[(size_t)ty * tiles_x + tx].push_back({zkey, t});
        }
        std::vector<unsigned long long> tile_keys(TILE * TILE);
        for (uint32_t tile = 0; tile < tiles_x * tiles_y; ++tile) {
            std::vector<std::pair<uint32_t, uint32_t>> &bin = bins[tile];
            if (bin.empty()) continue;
            std::stable_sort(bin.begin(), bin.end(), [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) { return a.first < b.first; });
            const uint32_t ox = (tile % tiles_x) * TILE, oy = y0 + (tile / tiles_x) * TILE;
            for (uint32_t i = 0; i < TILE * TILE; ++i) tile_keys[i] = (ox + (i % TILE) < W && oy + (i / TILE) < y1) ? VIS_EMPTY : 0ull;
            uint32_t block_far[8];
            auto refresh = [&]() {
                for (uint32_t b = 0; b < 8u; ++b) {
                    uint32_t m = 0u;
                    for (uint32_t yy = 0; yy < 8u; ++yy)
                        for (uint32_t xx = 0; xx < 16u; ++xx) m = std::max(m, (uint32_t)(tile_keys[((b >> 1) * 8u + yy) * TILE + (b & 1u) * 16u + xx] >> 32));
                    block_far[b] = m;
                }
            };
            refresh();
            for (const std::pair<uint32_t, uint32_t> &e : bin) {
                const uint32_t t = e.second;
                const float4 v0 = rv[tris[10 * (size_t)t]], v1 = rv[tris[10 * (size_t)t + 1]], v2 = rv[tris[10 * (size_t)t + 2]];
                const BBox bb = bounding_box(v0, v1, v2, vw);
                const uint32_t rx0 = max(bb.x0, ox), ry0 = max(bb.y0, oy), rx1 = min(bb.x1, ox + TILE - 1u), ry1 = min(bb.y1, oy + TILE - 1u);
                if (rx0 > rx1 || ry0 > ry1) continue;
                stage_item(*stg, 0u, t, 0u, v0, v1, v2, rx0, ry0, rx1, ry1, ox, oy);
                uint32_t keep = 0u;
                for (uint32_t b = 0; b < 8u; ++b)
                    if (((stg->w[23][0] >> b) & 1u) && !block_behind(b, block_far[b], rx0, ry0, rx1, ry1, ox, oy, exact::u2f(stg->w[27][0]), exact::u2f(stg->w[28][0]),
                                                                      exact::u2f(stg->w[29][0]), exact::u2f(stg->w[30][0])))
                        keep |= 1u << b;
                if (keep == 0u) continue;
                for (uint32_t lane = 0; lane < 32u; ++lane) raster_item<true, true>(*stg, 0u, lane, vw, nullptr, tile_keys.data(), true, nullptr, keep);
                refresh();
            }
            for (uint32_t i = 0; i < TILE * TILE; ++i) {
                const uint32_t x = ox + (i % TILE), y = oy + (i / TILE);
                if (tile_keys[i] != VIS_EMPTY && x < W && y < y1) {
                    unsigned long long &dst = vis[(size_t)(y - y0) * W + x];
                    if (tile_keys[i] < dst) dst = tile_keys[i];
                }
            }
        }
    }
#endif
    delete stg;

    // ---- k_prepare_tris (variant RAST_SHADE_PREP) ----
    std::vector<float4> prep;
    if (flags & EMU_PREP) {
        prep.resize((size_t)T * PREP_QUADS);
        for (uint64_t t = 0; t < T; ++t) prepare_triangle((uint32_t)t, sc, rv.data(), cn.data(), prep.data() + t * PREP_QUADS);
    }

    // ---- k_resolve_shade ----
    const bool pre = (flags & EMU_PRE_NORMALS) != 0, flat = (flags & EMU_FLAT_FACE) != 0;
    for (uint32_t y = y0; y < y1; ++y)
        for (uint32_t x = 0; x < W; ++x) {
            const size_t i = (size_t)(y - y0) * W + x;
            Shaded px;
            px.r = px.g = px.b = 0u; px.depth = 1.0f;
            uint32_t id = INVALID_TRI;
            if (vis[i] != VIS_EMPTY) {
                id = (uint32_t)vis[i];
                if (flags & EMU_PREP) px = shade_pixel_prep(id, x, y, sc, prep.data(), wind_clockwise != 0, lt, ld.data());
                else if (flat) px = shade_pixel<false, true>(id, x, y, sc, rv.data(), cn.data(), modelview, wind_clockwise != 0, lt, ld.data());
                else if (pre) px = shade_pixel<true, false>(id, x, y, sc, rv.data(), cn.data(), normal_m, wind_clockwise != 0, lt, ld.data());
                else px = shade_pixel<false, false>(id, x, y, sc, rv.data(), cn.data(), normal_m, wind_clockwise != 0, lt, ld.data());
            }
            rgb[i] = (uint8_t)px.r; rgb[i + P] = (uint8_t)px.g; rgb[i + 2 * P] = (uint8_t)px.b;
            depth[i] = px.depth;
            tri_ids[i] = id;
        }
    return 0;
}
