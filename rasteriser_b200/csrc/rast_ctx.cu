// rast_ctx.cu -- context, buffer management, launch sequencing and the extern "C" rast_* ABI
// (include/rast.h).  Host orchestration of draw_frame (drawing.cpp:205-258): the reference builds
// matrices, runs six whole-array passes and loops the faces in order; here the host builds the
// same matrices (hostmath.h), uploads them, and launches clear -> vertex -> setup -> chunk raster ->
// resolve+shade for a whole batch of frames at once.  No CPU fallback exists: every entry point
// either runs the CUDA kernels or fails with an error code.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <emmintrin.h>

#include "../../include/rast.h"
#include "hostmath.h"
#include "kernels.cuh"

namespace {

thread_local std::string g_create_error;

// Frames in flight through one launch sequence (<= 256: 8-bit frame tag).  Device-pointer draws take big batches: a launch sequence has fixed
// costs (ten launches, the ramp and the tail of every grid, a persistent raster grid that needs several grabs per warp to balance) -- 120-frame
// 1080p calls on a B200: 32 frames per batch 1.908 ms, 64: 1.853, 120: 1.816; the 720-frame step of bench.py at 90 / 120 / 180 / 240 per batch:
// 70.5 / 71.0 / 71.6 / 71.9 k frames/s (32: 67.7 k).  Host-buffer draws keep 32: a batch is also the unit in which frames travel back, and the
// first delivery should not wait for 240 frames.
#ifndef RAST_MAX_BATCH
#define RAST_MAX_BATCH 240
#endif
constexpr uint32_t MAX_BATCH = RAST_MAX_BATCH, MAX_BATCH_HOST = 32;
#ifndef RAST_BATCH_GB
#define RAST_BATCH_GB 12
#endif
constexpr size_t BATCH_BYTES_BUDGET = (size_t)RAST_BATCH_GB << 30;  // per-batch device scratch budget
constexpr unsigned long long TILE_MODE_OVERDRAW = 8; // queued bbox area per pixel above which the next call bins by screen tile
constexpr size_t SPARSE_MIN_FRAME_BYTES = 4u << 20; // host-buffer draws: frames smaller than this are copied whole (RAST_SPARSE_MIN_BYTES)
constexpr uint32_t QUEUE_MIN = 1u << 23;           // work items (8 B each); grows on demand when a frame overflows it

struct DeviceBuffer {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, want ? want : 1);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct PinnedBuffer {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMallocHost(&p, want ? want : 1);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// A few host threads for the one piece of per-pixel host work the path has: writing the constant background of the
// caller's frame / depth buffers when only the covered rectangle of a frame crosses PCIe (finish_batch).  The caller's
// thread takes tasks too; run() returns when all are done.
class HostPool {
public:
    ~HostPool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (std::thread &t : threads_) t.join();
    }
    void start(unsigned helpers) {
        for (unsigned i = threads_.size(); i < helpers; ++i) threads_.emplace_back([this]() { worker(); });
    }
    void run(const std::function<void(size_t)> &fn, size_t n) {
        if (threads_.empty() || n <= 1) {
            for (size_t i = 0; i < n; ++i) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn; n_ = n; next_ = 0; active_ = (unsigned)threads_.size(); ++generation_;
        }
        cv_.notify_all();
        for (size_t i; (i = next_.fetch_add(1)) < n;) fn(i);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this]() { return active_ == 0; });
    }

private:
    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> l(m_);
            cv_.wait(l, [&]() { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            const std::function<void(size_t)> *fn = fn_;
            const size_t n = n_;
            l.unlock();
            for (size_t i; (i = next_.fetch_add(1)) < n;) (*fn)(i);
            l.lock();
            if (--active_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t n_ = 0;
    std::atomic<size_t> next_{0};
    unsigned active_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

} // namespace

struct rast_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    std::string error;
    uint64_t launches = 0;
    uint64_t d2h_bytes = 0; // bytes of frame / depth data copied to host buffers so far (rast_d2h_bytes)
    unsigned raster_grid = 148; // persistent grid of k_raster_chunks: SMs x resident CTAs per SM

    // scene
    rk::Scene scene{};
    DeviceBuffer d_pos, d_nrm, d_nrm4, d_uv, d_vidx, d_attr, d_mats, d_texels;
    bool have_mesh = false;
    bool mesh_materials_dirty = false; // material indices in d_attr still have to be clamped against n_materials
    uint32_t n_materials = 0;
    bool pre_normals = false;
    uint64_t out_plane_stride = 0; // rast_set_output_plane_stride (0 = the band's own pixel count)
    uint32_t out_frame_stride = 1; // rast_set_output_frame_stride (device-pointer draws: frame slots between consecutive frames)
    bool flat_face = false; // extension mode of the current call (rast_args.flat == RAST_FLAT_FACE)
    rk::LightTable light_table{}; // first PARAM_LIGHTS lights, passed to the shade kernel by value
    std::vector<rast_light> lights;

    // view
    uint32_t band_y0 = 0, band_y1 = 0; // 0,0 = whole frame

    // per-call / per-batch buffers
    DeviceBuffer d_queue, d_counters, d_aux, d_tiles, d_items;
    // per-call parameter blocks, alternating between calls (cs = call slot): the front passes of call k+1 -- parameter upload
    // included -- may run while the shade pass of call k still reads its own block
    DeviceBuffer d_frames[2], d_lights[2];
    cudaEvent_t ev_call_done[2] = {nullptr, nullptr};
    bool call_done_pending[2] = {false, false};
    int cs = 0, next_ps = 0;
    // What the shade pass of batch b reads while vertex / setup / raster of batch b+1 write it exists twice (pipeline slot =
    // batch parity): the front passes of the next batch run on a high-priority stream of their own while the shade pass of
    // this batch runs on the context's stream -- both are issue-bound at 66-76 % and fill each other's gaps (two contexts
    // on one GPU measured +13 %; RAST_OVERLAP=0 serialises).
    DeviceBuffer d_rv[2], d_cn[2], d_vis[2], d_flags[2]; // d_flags: one byte per 32 x 16 pixel tile and frame (kernels.cuh, tile flags), cleared with the keys
#if RAST_SHADE_PREP
    DeviceBuffer d_prep[2];   // prepared shading records of a batch (kernels.cuh, k_prepare_tris), by pipeline slot like rv / cn
    bool use_prep = false;    // this call: records fit (decided in draw_frames_impl)
    bool prep_enabled = true; // RAST_SHADE_PREP_RUNTIME=0 keeps the variant build on the gather path (A/B inside one library)
#endif
    cudaStream_t front_stream = nullptr;
    cudaEvent_t ev_raster[2] = {nullptr, nullptr}, ev_shade[2] = {nullptr, nullptr};
    bool shade_pending[2] = {false, false};
    bool overlap = true;
    bool setup_pipe = true;               // meshes of >= SETUP_PIPE_MIN_TRIANGLES triangles take the software-pipelined setup kernel (RAST_SETUP_PIPE=0: never)
    unsigned setup_pipe_grid = 148;       // its persistent grid: SMs x resident CTAs per SM (x RAST_SETUP_PIPE_WAVES)
    uint32_t shade_wt_min_tiles = 16384; // batches with at least this many 32 x 16 tiles take the one-warp-per-tile shade kernel (RAST_SHADE_WT_MIN_TILES)
    unsigned shade_wt_grid = 148;         // persistent flavour of it (RAST_SHADE_PERSIST): SMs x resident CTAs per SM
    DeviceBuffer d_shade_cursor;
    bool keep_visibility = false; // rast_set_keep_visibility: the last frame of a call keeps its keys (rast_read_triangle_ids / rast_get_stats)
    int last_ps = 0; // pipeline slot of the most recent batch (rast_read_triangle_ids / rast_get_stats)
    DeviceBuffer d_rgb[2], d_depth[2];
    PinnedBuffer h_frames, h_lights, h_status, h_spans[2];
    DeviceBuffer d_spans[2];
    bool sparse_copy = true;     // host-buffer draws copy only each frame's covered rectangle back (RAST_SPARSE_COPY=0: whole frames)
    bool sparse_now = true;      // this call: sparse_copy and frames big enough for it to pay (sparse_min_bytes)
    size_t sparse_min_bytes = SPARSE_MIN_FRAME_BYTES; // RAST_SPARSE_MIN_BYTES
    // rast_set_retained_outputs: the caller promises that the host buffers of a draw still hold what this context's previous
    // host-buffer draw wrote there.  `retained` remembers that previous draw (buffers, geometry, one rectangle per frame: what
    // is NOT the cleared background); finish_batch then resets only the part of the old rectangle the new one does not cover.
    struct Retained {
        const uint8_t *frames = nullptr;
        const float *depths = nullptr;
        uint32_t W = 0, rows = 0, y0 = 0;
        std::vector<uint32_t> ext; // per frame and row of the buffers: [xa, xb) = what is not the cleared background
        bool valid = false;
    } retained;
    // zero-copy delivery (k_deliver): this call's host buffers as the device sees them, or nullptr (copy-engine path)
    uint8_t *frames_mapped = nullptr;
    float *depths_mapped = nullptr;
    bool deliver_now = false;
    bool deliver_enabled = true; // RAST_DELIVER=0: always the copy engine
    bool retained_outputs = false; // the promise (off by default)
    bool retained_now = false;     // this call: the promise holds for these buffers
    unsigned host_threads = 0;   // helpers of the background fill (RAST_HOST_THREADS = total threads; default min(4, hardware / 2))
    HostPool pool;
    cudaEvent_t ev_params = nullptr, ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    bool copied_pending[2] = {false, false};
    uint32_t queue_cap = QUEUE_MIN;
    // raster schedule: 0 = chunk queue, 1 = screen-tile bins; "auto" follows the overdraw estimate of the previous call
    uint32_t tiny_max_pixels = rk::TINY_MIN_PIXELS;
    bool tiny_max_forced = false;
    int raster_mode_forced = -1; // -1 auto, 0 chunk, 1 tile (RAST_RASTER_MODE)
    bool tile_mode_next = false;
    uint32_t items_cap = 1u << 25; // item slots of the screen-tile bins (all tiles of a batch together)
    // visibility-buffer bookkeeping: slots [0, vis_clean_slots) of band size vis_clean_pixels hold VIS_EMPTY,
    // except vis_dirty_slot (the last frame of the previous call, kept for inspection)
    uint32_t vis_clean_slots[2] = {0, 0}, vis_clean_pixels[2] = {0, 0}, vis_clean_w[2] = {0, 0};
    int vis_dirty_slot[2] = {-1, -1};

    // most recent frame (for rast_read_triangle_ids / rast_depth_to_u8 / stats)
    rk::View last_view{};
    uint32_t last_slot_frame = 0; // index of the last frame inside d_vis / d_rv
    const float *last_depth_dev = nullptr;
    size_t last_frames_offset = 0; // index into d_frames of the last frame's params
    bool have_frame = false, have_visibility = false;
    uint64_t last_queue_count = 0;
    size_t call_frames = 0; // frames of the call being launched
    unsigned long long last_batch_pixels = 0; // pixels (all frames) of the last batch launched

    // which kernel flavour each pass of the most recent batch took (rast_last_schedule)
    const char *sched_setup = "", *sched_shade = "";
    bool sched_bins_requested = false;
    uint32_t sched_batch = 0; // frames per launch sequence of the most recent call
    std::string schedule_text;

    // profiling
    bool profiling = false;
    cudaEvent_t ev_pass[RAST_PASS_COUNT + 1] = {};
    float pass_ms[RAST_PASS_COUNT] = {};
};

namespace {

int fail(rast_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess) {
    if (ctx) {
        ctx->error = what;
        if (e != cudaSuccess) { ctx->error += ": "; ctx->error += cudaGetErrorString(e); }
    }
    return code;
}

#define RAST_CUDA(ctx, call)                                            \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fail(ctx, RAST_ECUDA, #call, e__); \
    } while (0)

inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

rk::View make_view(const rast_ctx *ctx, uint32_t W, uint32_t H) {
    rk::View v;
    v.W = W;
    v.H = H;
    v.y0 = 0;
    v.y1 = H;
    if (ctx->band_y1 > ctx->band_y0) {
        v.y0 = ctx->band_y0 < H ? ctx->band_y0 : H;
        v.y1 = ctx->band_y1 < H ? ctx->band_y1 : H;
    }
    v.band_pixels = W * (v.y1 - v.y0);
    v.out_plane = v.band_pixels;
    v.out_frame_stride = 1;
    return v;
}

// host side of draw_frame for one frame: matrices (drawing.cpp:222-229, geometry.cpp:101)
void fill_frame_params(const rast_args &a, rk::FrameParams &fp) {
    using namespace hostmath;
    const float view_disp[3] = {0.f, 0.f, -3.f}, zero3[3] = {0.f, 0.f, 0.f};
    const Mat4 model = transformation_matrix(a.scale, a.displacement, a.tait_bryan_angles);
    const Mat4 view = transformation_matrix(1.f, view_disp, zero3);
    const Mat4 modelview = mul(view, model);
    camera_matrix(modelview, a.aspect_ratio).store(fp.camera);
    transpose(inverse(modelview)).store(fp.normal_m);
    fp.wind_clockwise = a.wind_clockwise ? 1u : 0u;
    fp.flat_face = a.flat == RAST_FLAT_FACE ? 1u : 0u;
    fp.pad[0] = fp.pad[1] = 0u;
    modelview.store(fp.modelview);
}

uint32_t batch_capacity(const rast_ctx *ctx, const rk::View &vw, bool device_ptrs) {
    const size_t per_frame = (size_t)vw.band_pixels * (8 + 2 * 7) + (size_t)ctx->scene.V * 16 + 256;
    size_t b = BATCH_BYTES_BUDGET / (per_frame ? per_frame : 1);
    const size_t cap = device_ptrs ? MAX_BATCH : std::min(MAX_BATCH, MAX_BATCH_HOST);
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (uint32_t)b;
}

// Launch the five passes for frames [first, first+count) of the uploaded parameter block.
// ps = pipeline slot (which copy of rv / cn / vis); two_streams: the shade pass goes to ctx->front_stream, ordered after this
// batch's raster pass by an event.  *done_stream receives the stream on which the batch's last kernel was launched.
int launch_batch(rast_ctx *ctx, const rk::View &vw, size_t first, uint32_t count, uint8_t *rgb_dev, float *depth_dev, uint32_t keep_frame,
                 uint32_t *spans_dev, int ps, bool two_streams, cudaStream_t *done_stream) {
    rk::Batch bt;
    bt.tile_min_area = ~0ull;
    bt.spans = spans_dev;
    bt.frames = ctx->d_frames[ctx->cs].as<rk::FrameParams>() + first;
    bt.n_frames = count;
    bt.rv = ctx->d_rv[ps].as<float4>();
    bt.cn = ctx->pre_normals ? ctx->d_cn[ps].as<float4>() : nullptr;
    bt.vis = ctx->d_vis[ps].as<unsigned long long>();
    bt.tile_flags = ctx->d_flags[ps].as<uint8_t>();
    const size_t flags_per_frame = (size_t)rk::flag_tiles_x(vw) * rk::flag_tiles_y(vw);
    // the tile leaves the shade pass in 16-byte stores when every plane row segment is 16-byte aligned
    const bool wide = (vw.W % 16u == 0u) && (vw.out_plane % 16u == 0u) && (((uintptr_t)rgb_dev & 15u) == 0u) && (((uintptr_t)depth_dev & 15u) == 0u);
#if RAST_SHADE_PREP
    bt.prep = (ctx->use_prep && bt.cn) ? ctx->d_prep[ps].as<float4>() : nullptr;
#endif
    bt.queue = ctx->d_queue.as<uint2>();
    bt.queue_cap = ctx->queue_cap;
    bt.tiny_max_pixels = ctx->tiny_max_pixels;
    bt.counters = ctx->d_counters.as<unsigned long long>();
    const rk::Scene &sc = ctx->scene;
    cudaStream_t st = two_streams ? ctx->front_stream : ctx->stream;
    const bool prof = ctx->profiling;
    const size_t n_vis = (size_t)count * vw.band_pixels;

    if (ctx->shade_pending[ps]) { // the shade pass that last read this slot's rv / cn / vis must be done before they are rewritten
        RAST_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_shade[ps], 0));
        ctx->shade_pending[ps] = false;
    }
    if (prof) cudaEventRecord(ctx->ev_pass[0], st);
    // Visibility slots known to be empty (handed back by the previous shade pass) are not cleared again.
    uint32_t n_clear_launches = 0;
    if (ctx->vis_clean_pixels[ps] != vw.band_pixels || ctx->vis_clean_w[ps] != vw.W) { ctx->vis_clean_slots[ps] = 0; ctx->vis_dirty_slot[ps] = -1; } // (the tile-flag layout follows W)
    if (ctx->vis_dirty_slot[ps] >= 0 && (uint32_t)ctx->vis_dirty_slot[ps] < ctx->vis_clean_slots[ps]) { // only one dirty slot is tracked: settle it now
        rk::k_clear<<<grid_for(((size_t)vw.band_pixels + 1) / 2, 256), 256, 0, st>>>(bt.vis + (size_t)ctx->vis_dirty_slot[ps] * vw.band_pixels, vw.band_pixels);
        RAST_CUDA(ctx, cudaMemsetAsync(bt.tile_flags + (size_t)ctx->vis_dirty_slot[ps] * flags_per_frame, 0, flags_per_frame, st));
        ctx->vis_dirty_slot[ps] = -1;
        n_clear_launches++;
    }
    if (count > ctx->vis_clean_slots[ps]) {
        const size_t from = (size_t)ctx->vis_clean_slots[ps] * vw.band_pixels;
        rk::k_clear<<<grid_for((n_vis - from + 1) / 2, 256), 256, 0, st>>>(bt.vis + from, n_vis - from);
        RAST_CUDA(ctx, cudaMemsetAsync(bt.tile_flags + (size_t)ctx->vis_clean_slots[ps] * flags_per_frame, 0, (count - ctx->vis_clean_slots[ps]) * flags_per_frame, st));
        if (ctx->vis_dirty_slot[ps] >= (int)ctx->vis_clean_slots[ps]) ctx->vis_dirty_slot[ps] = -1;
        ctx->vis_clean_slots[ps] = count;
        n_clear_launches++;
    }
    ctx->vis_clean_pixels[ps] = vw.band_pixels;
    ctx->vis_clean_w[ps] = vw.W;
    if (prof) cudaEventRecord(ctx->ev_pass[1], st);
    if (sc.V) rk::k_vertex<<<dim3(grid_for((size_t)sc.V + (bt.cn ? sc.Nn : 0u), 256), count), 256, 0, st>>>(sc, vw, bt);
#if RAST_SHADE_PREP
    if (bt.prep && sc.T) { rk::k_prepare_tris<<<dim3(grid_for(sc.T, 128), count), 128, 0, st>>>(sc, bt); ctx->launches++; }
#endif
    if (prof) cudaEventRecord(ctx->ev_pass[2], st);
    // Raster schedule of this batch: the bbox-anchored chunk queue (k_raster_chunks, any order, global atomicMin) for frames where few
    // fragments compete per pixel, screen-tile bins (k_raster_tiles: the tile's keys in shared memory, bin processed near to far, block
    // and item level depth rejection) at high overdraw.  RAST_RASTER_MODE=chunk|tile forces one.
    // "auto" bins the batch when the previous call's queued bbox area showed high overdraw (TILE_MODE_OVERDRAW); the raster kernels confirm it
    // for this batch on the device (tile_schedule_taken) and otherwise leave the batch to the chunk queue, which k_setup fills in either case
    const bool tile_mode = sc.T && (ctx->raster_mode_forced == 1 || (ctx->raster_mode_forced == -1 && ctx->tile_mode_next));
    rk::TileBins tb{};
    uint32_t n_tile_launches = 0;
    bt.tile_min_area = ~0ull;
    if (tile_mode) {
        tb.tiles_x = (vw.W + rk::TILE - 1) / rk::TILE;
        tb.tiles_y = (vw.y1 - vw.y0 + rk::TILE - 1) / rk::TILE;
        const size_t n_tiles = (size_t)count * tb.tiles_x * tb.tiles_y;
        // bins of fixed capacity, filled by k_setup itself: as many slots per tile as the item budget allows (8K frame: 32 400 tiles -> 1 035 slots for
        // bins of ~160); an overflowing bin sends the batch to the chunk queue and doubles the budget for the next call
        tb.cap = (uint32_t)std::max<size_t>(16, std::min<size_t>(4096, ctx->items_cap / (n_tiles ? n_tiles : 1)));
        RAST_CUDA(ctx, ctx->d_tiles.reserve(n_tiles * 4));
        RAST_CUDA(ctx, ctx->d_items.reserve(n_tiles * tb.cap * sizeof(uint2)));
        tb.fill = ctx->d_tiles.as<uint32_t>();
        tb.items = ctx->d_items.as<uint2>();
        RAST_CUDA(ctx, cudaMemsetAsync(tb.fill, 0, n_tiles * 4, st));
        bt.tile_min_area = ctx->raster_mode_forced == 1 ? 0ull : (unsigned long long)(TILE_MODE_OVERDRAW / 2) * vw.band_pixels * count;
    }
    ctx->sched_bins_requested = tile_mode;
    if (sc.T && vw.band_pixels) {
        if (tile_mode) { rk::k_setup<true><<<dim3(grid_for(sc.T, 256 * rk::SETUP_TRIS), count), 256, 0, st>>>(sc, vw, bt, tb); ctx->sched_setup = "k_setup<1,1>"; }
        else if (ctx->setup_pipe && sc.T >= rk::SETUP_PIPE_MIN_TRIANGLES) { rk::k_setup_pipe<false><<<dim3(std::min(grid_for(sc.T, 256), ctx->setup_pipe_grid), count), 256, 0, st>>>(sc, vw, bt, tb); ctx->sched_setup = "k_setup_pipe<0>"; }
        else if (rk::SETUP_TRIS == 1 && sc.T >= rk::SETUP_TRIS2_MIN_TRIANGLES) { rk::k_setup<false, 2><<<dim3(grid_for(sc.T, 256 * 2), count), 256, 0, st>>>(sc, vw, bt, tb); ctx->sched_setup = "k_setup<0,2>"; }
        else { rk::k_setup<false><<<dim3(grid_for(sc.T, 256 * rk::SETUP_TRIS), count), 256, 0, st>>>(sc, vw, bt, tb); ctx->sched_setup = "k_setup<0,1>"; }
    }
    if (prof) cudaEventRecord(ctx->ev_pass[3], st);
    if (sc.T && vw.band_pixels) rk::k_raster_chunks<<<ctx->raster_grid, rk::RASTER_WARPS * 32, 0, st>>>(sc, vw, bt); // returns at once when the bins are used
    if (tile_mode && vw.band_pixels) {
        rk::k_raster_tiles<<<dim3(tb.tiles_x * tb.tiles_y, count), rk::TILE_WARPS * 32, 0, st>>>(sc, vw, bt, tb);
        n_tile_launches += 1;
    }
    // queue statistics of the call's last batch travel back from here (overflow => the queue grows before the next call): the
    // next call's front passes reset the counters and may start before this call's shade pass has finished
    if (first + count == ctx->call_frames) RAST_CUDA(ctx, cudaMemcpyAsync(ctx->h_status.p, ctx->d_counters.p, 64, cudaMemcpyDeviceToHost, st));
    if (prof) cudaEventRecord(ctx->ev_pass[4], st);
    if (two_streams) { // the shade pass runs on the context's stream, after this batch's raster pass
        RAST_CUDA(ctx, cudaEventRecord(ctx->ev_raster[ps], st));
        st = ctx->stream;
        RAST_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_raster[ps], 0));
    }
    if (spans_dev) cudaMemsetAsync(spans_dev, 0xFF, (size_t)count * (vw.y1 - vw.y0) * 8, st);
    if (vw.band_pixels) {
        const rk::LightDev *lights = ctx->d_lights[ctx->cs].as<rk::LightDev>();
        const uint32_t rows = vw.y1 - vw.y0;
        const dim3 grid(grid_for(vw.W, rk::SHADE_TILE_W), grid_for(rows, rk::SHADE_TILE_H), count);
        const unsigned threads = rk::SHADE_WARPS * 32;
        const rk::LightTable &lt = ctx->light_table;
        // one warp per whole tile for big batches, four warps per tile (4 rows each) when the batch has few tiles (kernels.cuh)
        const uint32_t n_tiles = (uint32_t)(flags_per_frame * count);
        const bool warp_tiles = n_tiles >= ctx->shade_wt_min_tiles;
        ctx->sched_shade = warp_tiles ? "k_resolve_shade_wt" : "k_resolve_shade";
        unsigned int *cursor = nullptr;
        unsigned wt_grid = grid_for(n_tiles, rk::SHADE_WT_WARPS);
#if RAST_SHADE_PERSIST
        if (warp_tiles) {
            cursor = ctx->d_shade_cursor.as<unsigned int>() + ps;
            RAST_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, st));
            wt_grid = ctx->shade_wt_grid;
        }
#endif
#define RAST_SHADE_LAUNCH(PRE, FLAT, PREP)                                                                                                        \
        do {                                                                                                                                      \
            if (warp_tiles) {                                                                                                                     \
                if (wide) rk::k_resolve_shade_wt<true, PRE, FLAT, PREP><<<wt_grid, rk::SHADE_WT_WARPS * 32, 0, st>>>(sc, vw, bt, lt, lights, rgb_dev, depth_dev, keep_frame, n_tiles, cursor); \
                else rk::k_resolve_shade_wt<false, PRE, FLAT, PREP><<<wt_grid, rk::SHADE_WT_WARPS * 32, 0, st>>>(sc, vw, bt, lt, lights, rgb_dev, depth_dev, keep_frame, n_tiles, cursor);     \
            } else if (wide) rk::k_resolve_shade<true, PRE, FLAT, PREP><<<grid, threads, 0, st>>>(sc, vw, bt, lt, lights, rgb_dev, depth_dev, keep_frame); \
            else rk::k_resolve_shade<false, PRE, FLAT, PREP><<<grid, threads, 0, st>>>(sc, vw, bt, lt, lights, rgb_dev, depth_dev, keep_frame);     \
        } while (0)
        if (ctx->flat_face) RAST_SHADE_LAUNCH(false, true, false); // extension: face normals (never taken for reference-compatible arguments)
#if RAST_SHADE_PREP
        else if (bt.cn && bt.prep) RAST_SHADE_LAUNCH(true, false, true);
#endif
        else if (bt.cn) RAST_SHADE_LAUNCH(true, false, false);
        else RAST_SHADE_LAUNCH(false, false, false);
#undef RAST_SHADE_LAUNCH
    }
    if (prof) cudaEventRecord(ctx->ev_pass[5], st);
    ctx->launches += n_tile_launches + n_clear_launches + (sc.V ? 1 : 0) + ((sc.T && vw.band_pixels) ? 2 : 0) + (vw.band_pixels ? 1 : 0);
    if (keep_frame < count) ctx->vis_dirty_slot[ps] = (int)keep_frame; // that slot still holds its keys (rast_read_triangle_ids)
    if (two_streams) {
        RAST_CUDA(ctx, cudaEventRecord(ctx->ev_shade[ps], st));
        ctx->shade_pending[ps] = true;
    }
    ctx->last_ps = ps;
    *done_stream = st;
    RAST_CUDA(ctx, cudaGetLastError());
    if (prof) {
        RAST_CUDA(ctx, cudaEventSynchronize(ctx->ev_pass[5]));
        for (int p = 0; p < RAST_PASS_COUNT; ++p) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev_pass[p], ctx->ev_pass[p + 1]);
            ctx->pass_ms[p] += ms;
        }
    }
    return RAST_OK;
}

// Fill [p, p + bytes) with a repeated 4-byte pattern using non-temporal stores for the 16-byte aligned middle: the lines
// are written once and never read by this thread, so a regular store's read-for-ownership would double the DRAM traffic.
// p is 4-byte aligned when the pattern is not a single repeated byte.
inline void stream_fill(void *p, size_t bytes, uint32_t pattern) {
    unsigned char *b = static_cast<unsigned char *>(p), *e = b + bytes;
    const unsigned char pat[4] = {(unsigned char)pattern, (unsigned char)(pattern >> 8), (unsigned char)(pattern >> 16), (unsigned char)(pattern >> 24)};
    while (b < e && ((uintptr_t)b & 15u)) { *b = pat[(uintptr_t)b & 3u]; ++b; }
    const __m128i v = _mm_set1_epi32((int)pattern);
    for (; b + 64 <= e; b += 64) {
        _mm_stream_si128(reinterpret_cast<__m128i *>(b), v);
        _mm_stream_si128(reinterpret_cast<__m128i *>(b + 16), v);
        _mm_stream_si128(reinterpret_cast<__m128i *>(b + 32), v);
        _mm_stream_si128(reinterpret_cast<__m128i *>(b + 48), v);
    }
    for (; b + 16 <= e; b += 16) _mm_stream_si128(reinterpret_cast<__m128i *>(b), v);
    while (b < e) { *b = pat[(uintptr_t)b & 3u]; ++b; }
}

// Host-buffer draws: bring one finished batch back.  Whole frames are one contiguous copy each.  With sparse_copy only the
// rectangle the shade pass reported as covered crosses PCIe (cudaMemcpy2DAsync per plane: 1000 x 900 of a 1080p frame moves
// at 49 GB/s of payload, tools/micro/d2h_2d_bench.cu) and the host writes the constant rest of the caller's buffers itself --
// frame 0, depth 1.0f, what the clear of renderer.cpp:85-86 leaves where nothing was drawn -- on the pool's threads while
// the copies are in flight.  The result in host memory is the same bytes either way.
struct PendingBatch {
    bool valid = false;
    int slot = 0;
    uint32_t first = 0, count = 0;
    const uint8_t *rgb_dev = nullptr;
    const float *depth_dev = nullptr;
};

int finish_batch(rast_ctx *ctx, const PendingBatch &b, const rk::View &vw, uint8_t *frames, float *depths) {
    const size_t P = vw.band_pixels;
    const uint32_t W = vw.W, rows = vw.y1 - vw.y0;
    cudaStream_t cs = ctx->copy_stream;
    if (!ctx->sparse_now) {
        RAST_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_done[b.slot], 0));
        if (frames) RAST_CUDA(ctx, cudaMemcpyAsync(frames + (size_t)b.first * 3 * P, b.rgb_dev, (size_t)b.count * 3 * P, cudaMemcpyDeviceToHost, cs));
        if (depths) RAST_CUDA(ctx, cudaMemcpyAsync(depths + (size_t)b.first * P, b.depth_dev, (size_t)b.count * P * 4, cudaMemcpyDeviceToHost, cs));
        ctx->d2h_bytes += (size_t)b.count * P * ((frames ? 3u : 0u) + (depths ? 4u : 0u));
        RAST_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b.slot], cs));
        ctx->copied_pending[b.slot] = true;
        return RAST_OK;
    }
    RAST_CUDA(ctx, cudaEventSynchronize(ctx->ev_done[b.slot])); // kernels done, row spans in h_spans[slot]
    const uint32_t *sp = ctx->h_spans[b.slot].as<uint32_t>();
    constexpr uint32_t S = rk::BBOX_STRIPS;
    const uint32_t bpp = (frames ? 3u : 0u) + (depths ? 4u : 0u);
    // ext[(i * rows + y) * 2 .. +1] = [xa, xb): the part of row y of frame i that is DELIVERED from the device (k_deliver: the row's span;
    // copy engine: the strip rectangle the row lies in); everything else of the row is background and written by the host.  Widened
    // to 64-pixel columns, so that no cache line is shared between the two writers.
    std::vector<uint32_t> ext((size_t)b.count * rows * 2, 0u);
    auto span_of = [&](uint32_t i, uint32_t y, uint32_t &xa, uint32_t &xb) {
        const uint32_t *q = sp + ((size_t)i * rows + y) * 2;
        if (q[0] == 0xFFFFFFFFu) { xa = xb = 0u; return; }
        xa = q[0] & ~63u;
        xb = std::min(W, ((W - 1u - q[1]) | 63u) + 1u);
    };
    struct Rect { uint32_t x0, y0, x1, y1; bool empty; };
    std::vector<Rect> rects;               // copy-engine path: one rectangle per (frame, strip of tile rows)
    std::vector<uint8_t> whole(b.count, 0); // copy-engine path: the rectangles cover more than 75 % of the frame -- one contiguous copy is cheaper
    if (ctx->deliver_now) {
        unsigned long long px = 0;
        for (uint32_t i = 0; i < b.count; ++i)
            for (uint32_t y = 0; y < rows; ++y) {
                uint32_t xa, xb;
                span_of(i, y, xa, xb);
                ext[((size_t)i * rows + y) * 2] = xa; ext[((size_t)i * rows + y) * 2 + 1] = xb;
                px += xb - xa;
            }
        ctx->d2h_bytes += px * bpp; // (k_deliver for this batch was queued behind its shade pass when the batch was launched)
    } else {
        rects.assign((size_t)b.count * S, Rect{0u, 0u, 0u, 0u, true});
        const uint32_t tiles_y = rk::flag_tiles_y(vw);
        for (uint32_t i = 0; i < b.count; ++i) {
            for (uint32_t y = 0; y < rows; ++y) {
                uint32_t xa, xb;
                span_of(i, y, xa, xb);
                if (xa == xb) continue;
                Rect &r = rects[(size_t)i * S + rk::bbox_strip_of_tile_row(y / rk::SHADE_TILE_H, tiles_y)];
                if (r.empty) r = Rect{xa, y, xb - 1u, y, false};
                else { r.x0 = std::min(r.x0, xa); r.x1 = std::max(r.x1, xb - 1u); r.y1 = y; }
            }
            uint64_t area = 0;
            for (uint32_t k = 0; k < S; ++k) {
                const Rect &r = rects[(size_t)i * S + k];
                if (!r.empty) area += (uint64_t)(r.x1 - r.x0 + 1u) * (r.y1 - r.y0 + 1u);
            }
            whole[i] = area * 4u > (uint64_t)P * 3u;
            for (uint32_t k = 0; k < S; ++k) {
                const Rect &r = rects[(size_t)i * S + k];
                if (whole[i]) continue;
                if (!r.empty)
                    for (uint32_t y = r.y0; y <= r.y1; ++y) { ext[((size_t)i * rows + y) * 2] = r.x0; ext[((size_t)i * rows + y) * 2 + 1] = r.x1 + 1u; }
            }
            if (whole[i])
                for (uint32_t y = 0; y < rows; ++y) { ext[((size_t)i * rows + y) * 2] = 0u; ext[((size_t)i * rows + y) * 2 + 1] = W; }
        }
    }
    // retained outputs: what each of these frames' buffers held outside the cleared background before this call
    std::vector<uint32_t> old_ext;
    if (ctx->retained_now) old_ext.assign(ctx->retained.ext.begin() + (size_t)b.first * rows * 2, ctx->retained.ext.begin() + ((size_t)b.first + b.count) * rows * 2);
    if (ctx->retained_outputs) { // remember what these buffers will hold (for the next call that keeps the promise)
        std::vector<uint32_t> &keep = ctx->retained.ext;
        if (keep.size() < ((size_t)b.first + b.count) * rows * 2) keep.resize(((size_t)b.first + b.count) * rows * 2, 0u);
        std::copy(ext.begin(), ext.end(), keep.begin() + (size_t)b.first * rows * 2);
    }
    // copy-engine path: the three colour planes of a rectangle are ONE 3-D copy (planes are P bytes apart), the depth plane a 2-D copy
    std::atomic<int> copy_error{(int)cudaSuccess};
    std::atomic<unsigned long long> copied{0};
    auto issue_copies = [&](uint32_t i) {
        const size_t f = (size_t)b.first + i;
        auto ok = [&](cudaError_t e) { if (e != cudaSuccess) { int want = (int)cudaSuccess; copy_error.compare_exchange_strong(want, (int)e); } };
        if (whole[i]) {
            if (frames) ok(cudaMemcpyAsync(frames + f * 3 * P, b.rgb_dev + (size_t)i * 3 * P, 3 * P, cudaMemcpyDeviceToHost, cs));
            if (depths) ok(cudaMemcpyAsync(depths + f * P, b.depth_dev + (size_t)i * P, P * 4, cudaMemcpyDeviceToHost, cs));
            copied += P * bpp;
            return;
        }
        for (uint32_t k = 0; k < S; ++k) {
            const Rect &r = rects[(size_t)i * S + k];
            if (r.empty) continue;
            const size_t off = (size_t)r.y0 * W + r.x0, w = r.x1 - r.x0 + 1u, h = r.y1 - r.y0 + 1u;
            copied += w * h * bpp;
            if (frames) {
                cudaMemcpy3DParms c3{};
                c3.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t *>(b.rgb_dev) + (size_t)i * 3 * P, W, W, rows);
                c3.dstPtr = make_cudaPitchedPtr(frames + f * 3 * P, W, W, rows);
                c3.srcPos = c3.dstPos = make_cudaPos(r.x0, r.y0, 0);
                c3.extent = make_cudaExtent(w, h, 3);
                c3.kind = cudaMemcpyDeviceToHost;
                ok(cudaMemcpy3DAsync(&c3, cs));
            }
            if (depths)
                ok(cudaMemcpy2DAsync(depths + f * P + off, (size_t)W * 4, b.depth_dev + (size_t)i * P + off, (size_t)W * 4, w * 4, h, cudaMemcpyDeviceToHost, cs));
        }
    };
    // the background, in tasks of one plane x 64 rows; on the copy-engine path task 0 issues the copies meanwhile
    const uint32_t chunk = 64, chunks = (rows + chunk - 1) / chunk, planes = (frames ? 3u : 0u) + (depths ? 1u : 0u);
    if (planes == 0) return RAST_OK;
    ctx->pool.start(ctx->host_threads);
    const bool retained = !old_ext.empty();
    const bool ce = !ctx->deliver_now;
    const int device = ctx->device;
    const std::function<void(size_t)> fill = [&](size_t task) {
        if (ce) {
            if (task == 0) {
                cudaSetDevice(device); // (a pool thread has no current device of its own)
                for (uint32_t i = 0; i < b.count; ++i) issue_copies(i);
                return;
            }
            --task;
        }
        const uint32_t i = (uint32_t)(task / ((size_t)planes * chunks)), rest = (uint32_t)(task % ((size_t)planes * chunks));
        const uint32_t plane = rest / chunks, ya = (rest % chunks) * chunk, yb = std::min(rows, ya + chunk);
        const size_t f = (size_t)b.first + i;
        const bool is_depth = frames ? plane == 3u : true;
        auto fill_span = [&](uint32_t y, uint32_t xa, uint32_t xb) { // [xa, xb) of row y
            if (xa >= xb) return;
            if (is_depth) stream_fill(depths + f * P + (size_t)y * W + xa, (size_t)(xb - xa) * 4, 0x3F800000u /* 1.0f */);
            else stream_fill(frames + (f * 3 + plane) * P + (size_t)y * W + xa, xb - xa, 0u);
        };
        for (uint32_t y = ya; y < yb; ++y) {
            const uint32_t xa = ext[((size_t)i * rows + y) * 2], xb = ext[((size_t)i * rows + y) * 2 + 1];
            // what of row y must become background: everything outside the delivered extent, or -- retained outputs -- only what the previous draw left there
            uint32_t oa = 0u, ob = W;
            if (retained) { oa = old_ext[((size_t)i * rows + y) * 2]; ob = old_ext[((size_t)i * rows + y) * 2 + 1]; }
            if (oa >= ob) continue;
            if (xa >= xb) { fill_span(y, oa, ob); continue; }
            fill_span(y, oa, std::min(ob, xa));
            fill_span(y, std::max(oa, xb), ob);
        }
        _mm_sfence(); // the streaming stores of this task are globally visible before it counts as done
    };
    ctx->pool.run(fill, (ce ? 1u : 0u) + (size_t)b.count * planes * chunks);
    _mm_sfence();
    if (ce) {
        ctx->d2h_bytes += copied.load();
        if (copy_error.load() != (int)cudaSuccess) return fail(ctx, RAST_ECUDA, "device-to-host copy of a covered rectangle", (cudaError_t)copy_error.load());
        RAST_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b.slot], cs));
        ctx->copied_pending[b.slot] = true;
    }
    return RAST_OK;
}

// Shared implementation of every draw entry point.
int draw_frames_impl(rast_ctx *ctx, const rast_args *args, uint32_t n, uint8_t *frames, float *depths, bool device_ptrs) {
    if (!ctx) return RAST_EINVAL;
    if (!args || n == 0) return fail(ctx, RAST_EINVAL, "rast_draw: no frames");
    if (!ctx->have_mesh) return fail(ctx, RAST_ESTATE, "rast_draw: rast_upload_mesh has not been called");
    const uint32_t W = args[0].image_width, H = args[0].image_height;
    if (W == 0 || H == 0 || W > 65535u || H > 65535u) return fail(ctx, RAST_EINVAL, "rast_draw: image size out of range");
    if ((uint64_t)W * H > 0xFFFFFFFFull) return fail(ctx, RAST_EINVAL, "rast_draw: more than 2^32 pixels");
    for (uint32_t i = 1; i < n; ++i) {
        if (args[i].image_width != W || args[i].image_height != H) return fail(ctx, RAST_EINVAL, "rast_draw_frames: all frames must share one image size");
        if ((args[i].flat == RAST_FLAT_FACE) != (args[0].flat == RAST_FLAT_FACE)) return fail(ctx, RAST_EINVAL, "rast_draw_frames: all frames must share one flat mode");
    }
    ctx->flat_face = args[0].flat == RAST_FLAT_FACE;
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->scene.mats == nullptr) { // no materials uploaded: everything uses the white sentinel
        int rc = rast_upload_materials(ctx, nullptr, 0);
        if (rc != RAST_OK) return rc;
    }
    rk::View vw = make_view(ctx, W, H);
    if (vw.band_pixels == 0) return RAST_OK; // empty band: nothing to render or copy
    // P = pixels between output planes.  Device-pointer draws may write a band into a larger image (the full frame of a
    // sort-first gather target, possibly another GPU's memory): then the planes are the larger image's.
    if (device_ptrs && ctx->out_plane_stride) {
        if (ctx->out_plane_stride < vw.band_pixels) return fail(ctx, RAST_EINVAL, "rast_draw_frames: output plane stride smaller than the band");
        vw.out_plane = (uint32_t)ctx->out_plane_stride;
    }
    if (device_ptrs) vw.out_frame_stride = ctx->out_frame_stride;
    const size_t S = vw.out_frame_stride;
    const size_t P = vw.out_plane;
    const uint32_t B = batch_capacity(ctx, vw, device_ptrs);
    const uint32_t nb = n < B ? n : B;
    ctx->sched_batch = nb;

    // the overflow flag of the previous call tells whether the work queue must grow
    if (ctx->h_status.as<unsigned long long>()[2] != 0ull) {
        RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const unsigned long long wanted = ctx->h_status.as<unsigned long long>()[0];
        uint64_t cap = ctx->queue_cap;
        while (cap < wanted && cap < (1ull << 27)) cap *= 2;
        ctx->queue_cap = (uint32_t)cap;
        ctx->h_status.as<unsigned long long>()[2] = 0ull;
    }

    {   // overdraw estimate of the previous call's last batch: queued bbox area / pixels rendered
        const unsigned long long *hs = ctx->h_status.as<unsigned long long>();
        if (ctx->last_batch_pixels) ctx->tile_mode_next = hs[3] > (unsigned long long)TILE_MODE_OVERDRAW * ctx->last_batch_pixels;
        if (hs[7] != 0ull) { // the bins overflowed (that batch fell back to the chunk queue): grow them
            RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->items_cap < (1u << 29)) ctx->items_cap *= 2;
            ctx->h_status.as<unsigned long long>()[7] = 0ull;
        }
    }

    // buffers
    ctx->cs ^= 1;
    RAST_CUDA(ctx, ctx->d_frames[ctx->cs].reserve((size_t)n * sizeof(rk::FrameParams)));
    RAST_CUDA(ctx, ctx->d_lights[ctx->cs].reserve((ctx->lights.size() + 1) * sizeof(rk::LightDev)));
    // The front passes (parameter upload, clear, vertex, setup, raster) go to the high-priority front stream and the shade pass
    // stays on the context's stream: inside a call batch b+1's front passes overlap batch b's shade pass, and across calls
    // the next call's front passes overlap this call's shade pass (they depend on nothing the context's stream produces).
    const bool two_streams = ctx->overlap && !ctx->profiling;
    for (int ps = 0; ps < 2; ++ps) {
        RAST_CUDA(ctx, ctx->d_rv[ps].reserve((size_t)nb * ctx->scene.V * sizeof(float4)));
        if ((size_t)ctx->scene.Nn * 8 <= P) RAST_CUDA(ctx, ctx->d_cn[ps].reserve((size_t)nb * ctx->scene.Nn * sizeof(float4)));
#if RAST_SHADE_PREP
        // prepared records only where they are small beside the visibility buffer (160 B per triangle against 8 B per pixel)
        // and where the extra launch pays: several batches (it then overlaps the previous batch's shade pass) or a lot of pixels to shade
        // (measured on a B200: 120-frame 1080p calls +5 %, one 8K frame shade pass -12 %, but a single 640x480 frame 51 -> 55 us)
        ctx->use_prep = ctx->prep_enabled && !ctx->flat_face && ctx->scene.T > 0 && (size_t)ctx->scene.Nn * 8 <= P && (size_t)ctx->scene.T * rk::PREP_QUADS * 16 <= (size_t)P * 8 &&
                        (n > nb || (size_t)nb * P >= ((size_t)16 << 20));
        if (ctx->use_prep) RAST_CUDA(ctx, ctx->d_prep[ps].reserve((size_t)nb * ctx->scene.T * rk::PREP_QUADS * sizeof(float4)));
#endif
        const size_t flag_bytes = (size_t)nb * rk::flag_tiles_x(vw) * rk::flag_tiles_y(vw);
        if ((size_t)nb * P * 8 > ctx->d_vis[ps].bytes || flag_bytes > ctx->d_flags[ps].bytes) { ctx->vis_clean_slots[ps] = 0; ctx->vis_dirty_slot[ps] = -1; }
        RAST_CUDA(ctx, ctx->d_vis[ps].reserve((size_t)nb * P * 8));
        RAST_CUDA(ctx, ctx->d_flags[ps].reserve(flag_bytes));
    }
    // camera-space normals are precomputed per frame when there are few of them relative to the image;
    // for huge meshes the shade pass transforms only the normals of winning triangles instead
    ctx->pre_normals = (size_t)ctx->scene.Nn * 8 <= P;
    RAST_CUDA(ctx, ctx->d_queue.reserve((size_t)ctx->queue_cap * sizeof(uint2)));
    RAST_CUDA(ctx, ctx->d_counters.reserve(128));

    // the previous call's parameter upload must have left the pinned staging block
    RAST_CUDA(ctx, cudaEventSynchronize(ctx->ev_params));
    RAST_CUDA(ctx, ctx->h_frames.reserve((size_t)n * sizeof(rk::FrameParams)));
    RAST_CUDA(ctx, ctx->h_lights.reserve((ctx->lights.size() + 1) * sizeof(rk::LightDev)));
    rk::FrameParams *hp = ctx->h_frames.as<rk::FrameParams>();
    for (uint32_t i = 0; i < n; ++i) fill_frame_params(args[i], hp[i]);

    // transform_lights(view, lights) (drawing.cpp:238): view is the fixed translate(0,0,-3)
    {
        const float view_disp[3] = {0.f, 0.f, -3.f}, zero3[3] = {0.f, 0.f, 0.f};
        const hostmath::Mat4 view = hostmath::transformation_matrix(1.f, view_disp, zero3);
        rk::LightDev *hl = ctx->h_lights.as<rk::LightDev>();
        for (size_t l = 0; l < ctx->lights.size(); ++l) {
            rast_light &L = ctx->lights[l];
            hostmath::light_direction(view, L.direction, L.trans_dir);
            hl[l].ntx = -L.trans_dir[0];
            hl[l].nty = -L.trans_dir[1];
            hl[l].ntz = -L.trans_dir[2];
            hl[l].icr = L.intensity * L.colour[0]; // light.intensity * light.colour (shading.cpp:21)
            hl[l].icg = L.intensity * L.colour[1];
            hl[l].icb = L.intensity * L.colour[2];
            hl[l].pad0 = hl[l].pad1 = 0.f;
            if (l < rk::PARAM_LIGHTS) {
                ctx->light_table.a[l] = make_float4(hl[l].ntx, hl[l].nty, hl[l].ntz, hl[l].icr);
                ctx->light_table.c[l] = make_float2(hl[l].icg, hl[l].icb);
            }
        }
        ctx->light_table.n = (uint32_t)ctx->lights.size();
    }
    cudaStream_t up = two_streams ? ctx->front_stream : ctx->stream;
    if (ctx->call_done_pending[ctx->cs]) { // the call that last used this parameter block (two calls ago) must have finished shading
        RAST_CUDA(ctx, cudaStreamWaitEvent(up, ctx->ev_call_done[ctx->cs], 0));
        ctx->call_done_pending[ctx->cs] = false;
    }
    RAST_CUDA(ctx, cudaMemcpyAsync(ctx->d_frames[ctx->cs].p, hp, (size_t)n * sizeof(rk::FrameParams), cudaMemcpyHostToDevice, up));
    if (!ctx->lights.empty())
        RAST_CUDA(ctx, cudaMemcpyAsync(ctx->d_lights[ctx->cs].p, ctx->h_lights.p, ctx->lights.size() * sizeof(rk::LightDev), cudaMemcpyHostToDevice, up));
    RAST_CUDA(ctx, cudaEventRecord(ctx->ev_params, up));
    if (ctx->mesh_materials_dirty) {
        // Resolve material indices against the uploaded table; -1 / out of range -> sentinel (index n_materials).  Launched on the
        // stream that carries this call's front passes: k_prepare_tris (front stream) copies the resolved index into the prepared
        // records and the shade pass waits for the front passes, so both readers are ordered behind it whatever stream the caller
        // handed to rast_set_stream.  (Both uploads synchronise the context's stream, so no earlier shade pass still reads tri_rec.)
        if (ctx->scene.T) {
            rk::k_resolve_materials<<<grid_for(ctx->scene.T, 256), 256, 0, up>>>(ctx->d_attr.as<int4>(), ctx->scene.T, ctx->n_materials);
            ctx->launches++;
        }
        ctx->mesh_materials_dirty = false;
    }

    if (ctx->profiling) memset(ctx->pass_ms, 0, sizeof ctx->pass_ms);
    ctx->call_frames = n;
    // Small frames go back whole: the sparse copy waits on the host for the frame's covered rectangle, issues four 2-D copies and wakes the
    // fill threads -- more than the 40 us a 2 MB frame needs over PCIe (640x480 into page-locked buffers: 181 us per call sparse)
    ctx->sparse_now = ctx->sparse_copy && (size_t)vw.band_pixels * ((frames ? 3u : 0u) + (depths ? 4u : 0u)) >= ctx->sparse_min_bytes;
    // Zero-copy delivery: buffers that are page-locked AND mapped into the device's address space (rast_host_alloc, rast_host_register) receive
    // their covered row spans straight from a kernel (k_deliver); anything else goes through the copy engine, one rectangle per strip.
    ctx->deliver_now = false;
    ctx->frames_mapped = nullptr; ctx->depths_mapped = nullptr;
    if (!device_ptrs && ctx->sparse_now && ctx->deliver_enabled && vw.W % 16u == 0u && frames && (((uintptr_t)frames | (uintptr_t)depths) & 15u) == 0u) {
        auto mapped = [](const void *p) -> void * {
            cudaPointerAttributes a{};
            if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost && a.devicePointer != nullptr) return a.devicePointer;
            cudaGetLastError();
            return nullptr;
        };
        ctx->frames_mapped = static_cast<uint8_t *>(mapped(frames));
        ctx->depths_mapped = depths ? static_cast<float *>(mapped(depths)) : nullptr;
        ctx->deliver_now = ctx->frames_mapped != nullptr && (depths == nullptr || ctx->depths_mapped != nullptr);
    }
    if (!device_ptrs) {
        // retained outputs: the promise covers exactly the buffers of the previous host-buffer draw, same geometry, and only frames that draw wrote
        rast_ctx::Retained &rt = ctx->retained;
        ctx->retained_now = ctx->retained_outputs && ctx->sparse_now && rt.valid && rt.frames == frames && rt.depths == depths && rt.W == vw.W &&
                            rt.rows == vw.y1 - vw.y0 && rt.y0 == vw.y0 && (size_t)n * (vw.y1 - vw.y0) * 2 <= rt.ext.size();
        if (!ctx->retained_now) rt.ext.clear();
        rt.valid = false; // until this call has completed
        rt.frames = frames; rt.depths = depths; rt.W = vw.W; rt.rows = vw.y1 - vw.y0; rt.y0 = vw.y0;
    }

    int slot = 0, ps = ctx->next_ps;
    PendingBatch pending;
    for (uint32_t first = 0; first < n; first += nb, ps ^= 1) {
        const uint32_t count = (n - first) < nb ? (n - first) : nb;
        // outputs: the caller's device buffers, or the context's own (double-buffered when a D2H copy follows)
        uint8_t *rgb_dst;
        float *depth_dst = nullptr;
        if (device_ptrs && frames) {
            rgb_dst = frames + (size_t)first * S * 3 * P;
        } else {
            RAST_CUDA(ctx, ctx->d_rgb[slot].reserve((size_t)nb * 3 * P));
            rgb_dst = ctx->d_rgb[slot].as<uint8_t>();
        }
        if (device_ptrs && depths) {
            depth_dst = depths + (size_t)first * S * P;
        } else if (depths || n == 1) { // a single frame always keeps its depth (rast_depth_to_u8); a sequence only on request
            RAST_CUDA(ctx, ctx->d_depth[slot].reserve((size_t)nb * P * 4));
            depth_dst = ctx->d_depth[slot].as<float>();
        }
        if (ctx->copied_pending[slot]) { // the copy engine must be done reading this slot
            RAST_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[slot], 0));
            ctx->copied_pending[slot] = false;
        }
        // the last frame of the call keeps its visibility keys for rast_read_triangle_ids / rast_get_stats
        const uint32_t keep_frame = (ctx->keep_visibility && first + count == n) ? count - 1 : 0xFFFFFFFFu;
        uint32_t *spans_dev = nullptr;
        const size_t span_bytes = (size_t)(vw.y1 - vw.y0) * 8; // per frame
        if (!device_ptrs && ctx->sparse_now) {
            RAST_CUDA(ctx, ctx->d_spans[slot].reserve((size_t)nb * span_bytes));
            RAST_CUDA(ctx, ctx->h_spans[slot].reserve((size_t)nb * span_bytes));
            spans_dev = ctx->d_spans[slot].as<uint32_t>();
        }
        cudaStream_t done_stream = ctx->stream;
        int rc = launch_batch(ctx, vw, first, count, rgb_dst, depth_dst, keep_frame, spans_dev, ps, two_streams, &done_stream);
        if (rc != RAST_OK) return rc;

        ctx->last_view = vw;
        ctx->last_batch_pixels = (unsigned long long)count * vw.band_pixels;
        ctx->last_slot_frame = count - 1;
        ctx->last_frames_offset = first + count - 1;
        ctx->last_depth_dev = depth_dst ? depth_dst + (size_t)(count - 1) * S * P : nullptr;
        ctx->have_frame = true;
        ctx->have_visibility = keep_frame < count;

        if (!device_ptrs) {
            if (spans_dev) RAST_CUDA(ctx, cudaMemcpyAsync(ctx->h_spans[slot].p, spans_dev, (size_t)count * span_bytes, cudaMemcpyDeviceToHost, done_stream));
            RAST_CUDA(ctx, cudaEventRecord(ctx->ev_done[slot], done_stream));
            if (spans_dev && ctx->deliver_now) {
                // the covered spans leave for the caller's buffers as soon as the shade pass is done, behind nothing on the host
                RAST_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[slot], 0));
                const uint32_t rows_b = vw.y1 - vw.y0;
                rk::k_deliver<<<grid_for((size_t)count * rows_b * 32, 256), 256, 0, ctx->copy_stream>>>(spans_dev, rgb_dst, depths ? depth_dst : nullptr, ctx->frames_mapped + (size_t)first * 3 * P,
                                                                                                     ctx->depths_mapped ? ctx->depths_mapped + (size_t)first * P : nullptr, vw.W, rows_b, count);
                ctx->launches++;
                RAST_CUDA(ctx, cudaGetLastError());
                RAST_CUDA(ctx, cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream));
                ctx->copied_pending[slot] = true;
            }
            // the previous batch is brought back now, with this one already queued behind it on the GPU
            if (pending.valid) { rc = finish_batch(ctx, pending, vw, frames, depths); if (rc != RAST_OK) return rc; }
            pending.valid = true; pending.slot = slot; pending.first = first; pending.count = count;
            pending.rgb_dev = rgb_dst; pending.depth_dev = depth_dst;
            slot ^= 1;
        }
    }
    if (pending.valid) { int rc = finish_batch(ctx, pending, vw, frames, depths); if (rc != RAST_OK) return rc; }
    ctx->next_ps = ps;
    RAST_CUDA(ctx, cudaEventRecord(ctx->ev_call_done[ctx->cs], ctx->stream));
    ctx->call_done_pending[ctx->cs] = true;
    if (!device_ptrs) {
        RAST_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->copied_pending[0] = ctx->copied_pending[1] = false;
        ctx->last_queue_count = ctx->h_status.as<unsigned long long>()[0];
        ctx->retained.valid = ctx->retained_outputs && ctx->sparse_now; // (whole-frame copies keep no rectangles)
        ctx->retained_now = false;
    }
    return RAST_OK;
}

} // namespace

extern "C" {

const char *rast_version(void) { return "rasteriser_b200 0.1 (sm_100a)"; }

int rast_create(int device, rast_ctx **out) {
    if (!out) return RAST_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("rast_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return RAST_ECUDA;
    }
    if (device < 0 || device >= count) { g_create_error = "rast_create: device index out of range"; return RAST_EINVAL; }
    rast_ctx *ctx = new (std::nothrow) rast_ctx();
    if (!ctx) { g_create_error = "rast_create: out of host memory"; return RAST_ENOMEM; }
    ctx->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    {
        int least = 0, greatest = 0;
        ok = ok && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&ctx->front_stream, cudaStreamNonBlocking, greatest) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_params, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_raster[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_shade[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_call_done[i], cudaEventDisableTiming) == cudaSuccess;
    }
    for (int i = 0; i <= RAST_PASS_COUNT && ok; ++i) ok = ok && cudaEventCreate(&ctx->ev_pass[i]) == cudaSuccess;
    ok = ok && ctx->h_status.reserve(64) == cudaSuccess;
    if (ok) memset(ctx->h_status.p, 0, 64);
    if (ok) {
        int sms = 0, per_sm = 0;
        ok = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rk::k_raster_chunks, rk::RASTER_WARPS * 32, 0) == cudaSuccess;
        if (const char *e = getenv("RAST_RASTER_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));
        ctx->raster_grid = (unsigned)(sms > 0 ? sms : 148) * (unsigned)(per_sm > 0 ? per_sm : 1);
        int sp_per_sm = 0;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sp_per_sm, rk::k_setup_pipe<false>, 256, 0) == cudaSuccess;
        int waves = 1;
        if (const char *e = getenv("RAST_SETUP_PIPE_WAVES")) waves = std::max(1, atoi(e));
        ctx->setup_pipe_grid = (unsigned)(sms > 0 ? sms : 148) * (unsigned)(sp_per_sm > 0 ? sp_per_sm : 1) * (unsigned)waves;
        int wt_per_sm = 0;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wt_per_sm, rk::k_resolve_shade_wt<true, true, false, RAST_SHADE_PREP != 0>, rk::SHADE_WT_WARPS * 32, 0) == cudaSuccess;
        ctx->shade_wt_grid = (unsigned)(sms > 0 ? sms : 148) * (unsigned)(wt_per_sm > 0 ? wt_per_sm : 1);
        ok = ok && ctx->d_shade_cursor.reserve(64) == cudaSuccess;
    }
    if (!ok) {
        g_create_error = std::string("rast_create: CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError());
        rast_destroy(ctx);
        return RAST_ECUDA;
    }
    ctx->stream = ctx->own_stream;
    if (const char *e = getenv("RAST_SPARSE_COPY")) ctx->sparse_copy = atoi(e) != 0;
    if (const char *e = getenv("RAST_SPARSE_MIN_BYTES")) ctx->sparse_min_bytes = (size_t)atoll(e);
    if (const char *e = getenv("RAST_DELIVER")) ctx->deliver_enabled = atoi(e) != 0;
    if (const char *e = getenv("RAST_SETUP_PIPE")) ctx->setup_pipe = atoi(e) != 0;
    if (const char *e = getenv("RAST_SHADE_WT_MIN_TILES")) ctx->shade_wt_min_tiles = (uint32_t)atoll(e);
    if (const char *e = getenv("RAST_OVERLAP")) ctx->overlap = atoi(e) != 0;
#if RAST_SHADE_PREP
    if (const char *e = getenv("RAST_SHADE_PREP_RUNTIME")) ctx->prep_enabled = atoi(e) != 0;
#endif
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        ctx->host_threads = std::min(4u, std::max(1u, hw / 2u)) - 1u; // measured on the 16-core host: 4 / 8 / 16 threads -> 5.49 / 5.38 / 5.30 k frames/s (memory-bound)
        if (const char *e = getenv("RAST_HOST_THREADS")) ctx->host_threads = (unsigned)std::max(1, atoi(e)) - 1u;
    }
    if (const char *e = getenv("RAST_TINY_MAX")) { ctx->tiny_max_pixels = (uint32_t)atoi(e); ctx->tiny_max_forced = true; }
    if (const char *e = getenv("RAST_RASTER_MODE")) ctx->raster_mode_forced = !strcmp(e, "tile") ? 1 : (!strcmp(e, "chunk") ? 0 : -1);
    *out = ctx;
    return RAST_OK;
}

void rast_destroy(rast_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->front_stream) cudaStreamSynchronize(ctx->front_stream);
    DeviceBuffer *dev[] = {&ctx->d_pos, &ctx->d_nrm, &ctx->d_nrm4, &ctx->d_uv, &ctx->d_vidx, &ctx->d_attr, &ctx->d_mats, &ctx->d_texels, &ctx->d_frames[0], &ctx->d_frames[1], &ctx->d_lights[0], &ctx->d_lights[1],
                           &ctx->d_rv[0], &ctx->d_rv[1], &ctx->d_cn[0], &ctx->d_cn[1], &ctx->d_tiles, &ctx->d_items, &ctx->d_vis[0], &ctx->d_vis[1], &ctx->d_flags[0], &ctx->d_flags[1], &ctx->d_spans[0], &ctx->d_spans[1], &ctx->d_queue, &ctx->d_counters, &ctx->d_shade_cursor, &ctx->d_aux, &ctx->d_rgb[0], &ctx->d_rgb[1], &ctx->d_depth[0], &ctx->d_depth[1]};
    for (DeviceBuffer *b : dev) b->release();
#if RAST_SHADE_PREP
    ctx->d_prep[0].release();
    ctx->d_prep[1].release();
#endif
    ctx->h_frames.release();
    ctx->h_lights.release();
    ctx->h_status.release();
    ctx->h_spans[0].release();
    ctx->h_spans[1].release();
    if (ctx->ev_params) cudaEventDestroy(ctx->ev_params);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        if (ctx->ev_raster[i]) cudaEventDestroy(ctx->ev_raster[i]);
        if (ctx->ev_shade[i]) cudaEventDestroy(ctx->ev_shade[i]);
        if (ctx->ev_call_done[i]) cudaEventDestroy(ctx->ev_call_done[i]);
    }
    for (int i = 0; i <= RAST_PASS_COUNT; ++i)
        if (ctx->ev_pass[i]) cudaEventDestroy(ctx->ev_pass[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->front_stream) cudaStreamDestroy(ctx->front_stream);
    delete ctx;
}

const char *rast_last_error(const rast_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int rast_set_stream(rast_ctx *ctx, void *cuda_stream) {
    if (!ctx) return RAST_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return RAST_OK;
}

int rast_use_own_stream(rast_ctx *ctx) {
    if (!ctx) return RAST_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = ctx->own_stream;
    return RAST_OK;
}

int rast_upload_mesh(rast_ctx *ctx, const float *positions, uint32_t n_positions, const float *normals, uint32_t n_normals,
                     const float *uvs, uint32_t n_uvs, const int32_t *tris, uint64_t n_tris) {
    if (!ctx) return RAST_EINVAL;
    if ((n_positions && !positions) || (n_normals && !normals) || (n_uvs && !uvs) || (n_tris && !tris)) return fail(ctx, RAST_EINVAL, "rast_upload_mesh: null array");
    if (n_tris > 0xFFFFFFFEull) return fail(ctx, RAST_EINVAL, "rast_upload_mesh: triangle index must fit 32 bits");
    // The Triangle array goes to the device as it is; k_build_tri_records splits it into the SoA vertex indices and the
    // shade records and validates every index there (no per-triangle work on the host: 50 M triangles = 2 GB).
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RAST_CUDA(ctx, ctx->d_pos.reserve((size_t)n_positions * 12));
    RAST_CUDA(ctx, ctx->d_nrm.reserve(((size_t)n_normals + 1) * 12));
    RAST_CUDA(ctx, ctx->d_nrm4.reserve(((size_t)n_normals + 1) * 16));
    RAST_CUDA(ctx, ctx->d_uv.reserve(((size_t)n_uvs + 1) * 8));
    RAST_CUDA(ctx, ctx->d_vidx.reserve((size_t)n_tris * 3 * 4));
    RAST_CUDA(ctx, ctx->d_attr.reserve((size_t)n_tris * 3 * 16));
    RAST_CUDA(ctx, ctx->d_counters.reserve(128));
    // every copy is ordered on the context's stream (the caller's arrays are pageable: the calls return when the data
    // has left them), so the kernels that follow on that stream see the data whatever kind of stream it is
    RAST_CUDA(ctx, cudaMemsetAsync(ctx->d_nrm.as<float>() + (size_t)n_normals * 3, 0, 12, ctx->stream)); // sentinel: zero normal
    RAST_CUDA(ctx, cudaMemsetAsync(ctx->d_uv.as<float>() + (size_t)n_uvs * 2, 0, 8, ctx->stream));       // sentinel: uv (0,0)
    if (n_positions) RAST_CUDA(ctx, cudaMemcpyAsync(ctx->d_pos.p, positions, (size_t)n_positions * 12, cudaMemcpyHostToDevice, ctx->stream));
    if (n_normals) RAST_CUDA(ctx, cudaMemcpyAsync(ctx->d_nrm.p, normals, (size_t)n_normals * 12, cudaMemcpyHostToDevice, ctx->stream));
    if (n_uvs) RAST_CUDA(ctx, cudaMemcpyAsync(ctx->d_uv.p, uvs, (size_t)n_uvs * 8, cudaMemcpyHostToDevice, ctx->stream));
    rk::k_pad_normals<<<grid_for((size_t)n_normals + 1, 256), 256, 0, ctx->stream>>>(ctx->d_nrm.as<float>(), ctx->d_nrm4.as<float4>(), n_normals + 1u);
    ++ctx->launches;
    if (n_tris) {
        DeviceBuffer raw; // freed on return
        struct Release { DeviceBuffer &b; ~Release() { b.release(); } } release_raw{raw};
        RAST_CUDA(ctx, raw.reserve((size_t)n_tris * 40));
        RAST_CUDA(ctx, cudaMemcpyAsync(raw.p, tris, (size_t)n_tris * 40, cudaMemcpyHostToDevice, ctx->stream));
        unsigned long long *first_error = ctx->d_counters.as<unsigned long long>() + 15; // beyond the per-batch counters
        RAST_CUDA(ctx, cudaMemsetAsync(first_error, 0xFF, 8, ctx->stream));
        rk::k_build_tri_records<<<grid_for(n_tris, 256), 256, 0, ctx->stream>>>(raw.as<int>(), n_tris, n_positions, n_normals, n_uvs, ctx->d_vidx.as<int>(),
                                                                                 ctx->d_attr.as<int4>(), first_error);
        ++ctx->launches;
        unsigned long long err = ~0ull;
        RAST_CUDA(ctx, cudaMemcpyAsync(&err, first_error, 8, cudaMemcpyDeviceToHost, ctx->stream));
        RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (err != ~0ull) {
            ctx->have_mesh = false;
            switch (err & 3ull) {
                case rk::UPLOAD_ERR_VERTEX: return fail(ctx, RAST_EINVAL, "rast_upload_mesh: vertex index out of range");
                case rk::UPLOAD_ERR_NORMAL: return fail(ctx, RAST_EINVAL, "rast_upload_mesh: normal index out of range");
                default: return fail(ctx, RAST_EINVAL, "rast_upload_mesh: uv index out of range");
            }
        }
    }
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rk::Scene &s = ctx->scene;
    s.pos = ctx->d_pos.as<float>();
    s.nrm = ctx->d_nrm.as<float>();
    s.nrm4 = ctx->d_nrm4.as<float4>();
    s.uv = ctx->d_uv.as<float2>();
    s.vidx0 = ctx->d_vidx.as<int>();
    s.vidx1 = s.vidx0 + n_tris;
    s.vidx2 = s.vidx1 + n_tris;
    s.tri_rec = ctx->d_attr.as<int4>();
    s.V = n_positions;
    s.Nn = n_normals + 1;
    s.Nuv = n_uvs + 1;
    s.T = n_tris;
    ctx->mesh_materials_dirty = true;
    if (!ctx->tiny_max_forced) { // few triangles: keep bboxes parallel (queue them); millions: the setup thread rasterises more itself
        const uint64_t t = n_tris / 16384u;
        ctx->tiny_max_pixels = (uint32_t)(t < rk::TINY_MIN_PIXELS ? rk::TINY_MIN_PIXELS : (t > rk::TINY_MAX_PIXELS ? rk::TINY_MAX_PIXELS : t));
    }
    ctx->have_mesh = true;
    ctx->have_frame = false;
    return RAST_OK;
}

int rast_upload_materials(rast_ctx *ctx, const rast_material *materials, uint32_t n_materials) {
    if (!ctx) return RAST_EINVAL;
    if (n_materials && !materials) return fail(ctx, RAST_EINVAL, "rast_upload_materials: null array");
    std::vector<rk::MaterialDev> md((size_t)n_materials + 1);
    size_t texel_total = 0;
    {   // sentinel for material index -1 / out of range: untextured white (renderer.cpp:49)
        rk::MaterialDev &d = md[n_materials];
        d.kd[0] = d.kd[1] = d.kd[2] = 1.f;
        d.has_texture = 0; d.tex_w = d.tex_h = 0; d.texel_offset = 0;
    }
    for (uint32_t i = 0; i < n_materials; ++i) {
        const rast_material &m = materials[i];
        md[i].kd[0] = m.kd[0]; md[i].kd[1] = m.kd[1]; md[i].kd[2] = m.kd[2];
        md[i].has_texture = m.has_texture ? (1 | (m.has_texture & RAST_TEXTURE_MODULATE_KD)) : 0;
        md[i].tex_w = m.tex_w; md[i].tex_h = m.tex_h;
        md[i].texel_offset = (long long)texel_total;
        if (m.has_texture) {
            // CImg::linear_atXY throws on an empty image (CImg.h:13467-13470)
            if (!m.texels || m.tex_w <= 0 || m.tex_h <= 0) return fail(ctx, RAST_EINVAL, "rast_upload_materials: textured material without texels");
            texel_total += (size_t)m.tex_w * m.tex_h;
        }
    }
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RAST_CUDA(ctx, ctx->d_mats.reserve(md.size() * sizeof(rk::MaterialDev)));
    RAST_CUDA(ctx, ctx->d_texels.reserve(texel_total * sizeof(float4)));
    RAST_CUDA(ctx, cudaMemcpy(ctx->d_mats.p, md.data(), md.size() * sizeof(rk::MaterialDev), cudaMemcpyHostToDevice));
    for (uint32_t i = 0; i < n_materials; ++i)
        if (materials[i].has_texture) { // planar [3][h][w] (CImg layout) -> interleaved float4 per texel
            const size_t n = (size_t)materials[i].tex_w * materials[i].tex_h;
            std::vector<float4> inter;
            try { inter.resize(n); } catch (...) { return fail(ctx, RAST_ENOMEM, "rast_upload_materials: out of host memory"); }
            const float *t = materials[i].texels;
            for (size_t k = 0; k < n; ++k) inter[k] = make_float4(t[k], t[n + k], t[2 * n + k], 0.f);
            RAST_CUDA(ctx, cudaMemcpy(ctx->d_texels.as<float4>() + md[i].texel_offset, inter.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
        }
    ctx->scene.mats = ctx->d_mats.as<rk::MaterialDev>();
    ctx->scene.texels = ctx->d_texels.as<float4>();
    ctx->scene.M = n_materials + 1;
    ctx->n_materials = n_materials;
    ctx->mesh_materials_dirty = true;
    return RAST_OK;
}

int rast_set_lights(rast_ctx *ctx, const rast_light *lights, uint32_t n_lights) {
    if (!ctx) return RAST_EINVAL;
    if (n_lights && !lights) return fail(ctx, RAST_EINVAL, "rast_set_lights: null array");
    try { ctx->lights.assign(lights, lights + n_lights); } catch (...) { return fail(ctx, RAST_ENOMEM, "rast_set_lights: out of host memory"); }
    return RAST_OK;
}

void rast_frame_matrices(const rast_args *args, float modelview[16], float camera[16], float normal_matrix[16], float view[16]) {
    using namespace hostmath;
    const float view_disp[3] = {0.f, 0.f, -3.f}, zero3[3] = {0.f, 0.f, 0.f};
    const Mat4 model = transformation_matrix(args->scale, args->displacement, args->tait_bryan_angles);
    const Mat4 v = transformation_matrix(1.f, view_disp, zero3);
    const Mat4 mv = mul(v, model);
    if (modelview) mv.store(modelview);
    if (camera) camera_matrix(mv, args->aspect_ratio).store(camera);
    if (normal_matrix) transpose(inverse(mv)).store(normal_matrix);
    if (view) v.store(view);
}

void rast_transform_lights(const float view[16], rast_light *lights, uint32_t n_lights) {
    const hostmath::Mat4 v = hostmath::Mat4::load(view);
    for (uint32_t i = 0; i < n_lights; ++i) hostmath::light_direction(v, lights[i].direction, lights[i].trans_dir);
}

float rast_spin_angle(float ry0, uint32_t k, uint32_t n_frames) { return ry0 + (float)k * (6.2831853f / (float)n_frames); }

uint64_t rast_d2h_bytes(rast_ctx *ctx) { return ctx ? ctx->d2h_bytes : 0; }

int rast_set_output_plane_stride(rast_ctx *ctx, uint64_t pixels) {
    if (!ctx) return RAST_EINVAL;
    if (pixels > 0xFFFFFFFFull) return fail(ctx, RAST_EINVAL, "rast_set_output_plane_stride: more than 2^32 pixels");
    ctx->out_plane_stride = pixels;
    return RAST_OK;
}

int rast_set_output_frame_stride(rast_ctx *ctx, uint32_t frames) {
    if (!ctx) return RAST_EINVAL;
    if (frames == 0) return fail(ctx, RAST_EINVAL, "rast_set_output_frame_stride: stride must be at least 1");
    ctx->out_frame_stride = frames;
    return RAST_OK;
}

uint64_t rast_fnv1a64(const void *data, uint64_t bytes) {
    uint64_t h = 14695981039346656037ull;
    const unsigned char *b = static_cast<const unsigned char *>(data);
    for (uint64_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// ---- device memory shared between the processes of one node (one process per GPU) -----------------------------
void *rast_device_alloc(rast_ctx *ctx, uint64_t bytes) {
    if (!ctx) return nullptr;
    void *p = nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { fail(ctx, RAST_ENOMEM, "rast_device_alloc", cudaGetLastError()); return nullptr; }
    return p;
}

int rast_device_free(rast_ctx *ctx, void *device_ptr) {
    if (!ctx) return RAST_EINVAL;
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    RAST_CUDA(ctx, cudaFree(device_ptr));
    return RAST_OK;
}

int rast_device_read(rast_ctx *ctx, void *host_dst, const void *device_src, uint64_t bytes) {
    if (!ctx) return RAST_EINVAL;
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    RAST_CUDA(ctx, cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RAST_OK;
}

int rast_ipc_export(rast_ctx *ctx, void *device_ptr, unsigned char handle[RAST_IPC_HANDLE_BYTES]) {
    if (!ctx || !handle) return RAST_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == RAST_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    RAST_CUDA(ctx, cudaIpcGetMemHandle(&h, device_ptr));
    memcpy(handle, &h, sizeof h);
    return RAST_OK;
}

int rast_ipc_open(rast_ctx *ctx, const unsigned char handle[RAST_IPC_HANDLE_BYTES], void **device_ptr) {
    if (!ctx || !handle || !device_ptr) return RAST_EINVAL;
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    RAST_CUDA(ctx, cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess)); // maps the peer's memory (NVLink P2P)
    return RAST_OK;
}

int rast_ipc_close(rast_ctx *ctx, void *device_ptr) {
    if (!ctx) return RAST_EINVAL;
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RAST_CUDA(ctx, cudaIpcCloseMemHandle(device_ptr));
    return RAST_OK;
}

int rast_set_band(rast_ctx *ctx, uint32_t y0, uint32_t y1) {
    if (!ctx) return RAST_EINVAL;
    if (y1 < y0) return fail(ctx, RAST_EINVAL, "rast_set_band: y1 < y0");
    ctx->band_y0 = y0;
    ctx->band_y1 = y1;
    return RAST_OK;
}

int rast_draw_frame(rast_ctx *ctx, const rast_args *args, uint8_t *frame, float *depth, rast_light *lights_out) {
    if (!ctx) return RAST_EINVAL;
    if (!frame) return fail(ctx, RAST_EINVAL, "rast_draw_frame: frame buffer is null");
    int rc = draw_frames_impl(ctx, args, 1, frame, depth, false);
    if (rc == RAST_OK && lights_out)
        for (size_t l = 0; l < ctx->lights.size(); ++l) memcpy(lights_out[l].trans_dir, ctx->lights[l].trans_dir, sizeof(float) * 3);
    return rc;
}

int rast_draw_frame_device(rast_ctx *ctx, const rast_args *args, uint8_t *frame_dev, float *depth_dev) {
    return draw_frames_impl(ctx, args, 1, frame_dev, depth_dev, true);
}

int rast_draw_frames(rast_ctx *ctx, const rast_args *args, uint32_t n, uint8_t *frames, float *depths, int device_ptrs) {
    if (ctx && !device_ptrs && !frames) return fail(ctx, RAST_EINVAL, "rast_draw_frames: frames buffer is null");
    return draw_frames_impl(ctx, args, n, frames, depths, device_ptrs != 0);
}

int rast_sync(rast_ctx *ctx) {
    if (!ctx) return RAST_EINVAL;
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    ctx->last_queue_count = ctx->h_status.as<unsigned long long>()[0];
    return RAST_OK;
}

int rast_read_triangle_ids(rast_ctx *ctx, uint32_t *tri_ids) {
    if (!ctx || !tri_ids) return RAST_EINVAL;
    if (!ctx->have_frame) return fail(ctx, RAST_ESTATE, "rast_read_triangle_ids: no frame has been drawn");
    if (!ctx->have_visibility) return fail(ctx, RAST_ESTATE, "rast_read_triangle_ids: the last frame's visibility buffer was not kept (rast_set_keep_visibility)");
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t P = ctx->last_view.band_pixels;
    RAST_CUDA(ctx, ctx->d_aux.reserve((size_t)P * 4 + 64));
    const unsigned long long *vis = ctx->d_vis[ctx->last_ps].as<unsigned long long>() + (size_t)ctx->last_slot_frame * P;
    rk::k_extract_tri_ids<<<grid_for(P, 256), 256, 0, ctx->stream>>>(vis, ctx->d_aux.as<uint32_t>(), P);
    ctx->launches++;
    RAST_CUDA(ctx, cudaMemcpyAsync(tri_ids, ctx->d_aux.p, (size_t)P * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RAST_OK;
}

int rast_depth_to_u8(rast_ctx *ctx, uint8_t *out) {
    if (!ctx || !out) return RAST_EINVAL;
    if (!ctx->have_frame || !ctx->last_depth_dev) return fail(ctx, RAST_ESTATE, "rast_depth_to_u8: the last frame was drawn without a depth output");
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t P = ctx->last_view.band_pixels;
    RAST_CUDA(ctx, ctx->d_aux.reserve((size_t)P + 64));
    uint32_t *minmax = ctx->d_aux.as<uint32_t>();
    uint8_t *u8 = ctx->d_aux.as<uint8_t>() + 64;
    const uint32_t init[2] = {0xFFFFFFFFu, 0u};
    RAST_CUDA(ctx, cudaMemcpyAsync(minmax, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    rk::k_depth_minmax<<<148 * 4, 256, 0, ctx->stream>>>(ctx->last_depth_dev, P, minmax);
    rk::k_depth_to_u8<<<grid_for(P, 256), 256, 0, ctx->stream>>>(ctx->last_depth_dev, P, minmax, u8);
    ctx->launches += 2;
    RAST_CUDA(ctx, cudaMemcpyAsync(out, u8, P, cudaMemcpyDeviceToHost, ctx->stream));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RAST_OK;
}

int rast_get_stats(rast_ctx *ctx, rast_stats *out) {
    if (!ctx || !out) return RAST_EINVAL;
    if (!ctx->have_frame) return fail(ctx, RAST_ESTATE, "rast_get_stats: no frame has been drawn");
    if (!ctx->have_visibility) return fail(ctx, RAST_ESTATE, "rast_get_stats: the last frame's visibility buffer was not kept (rast_set_keep_visibility)");
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    RAST_CUDA(ctx, ctx->d_aux.reserve(64));
    unsigned long long *cnt = ctx->d_aux.as<unsigned long long>();
    RAST_CUDA(ctx, cudaMemsetAsync(cnt, 0, 16, ctx->stream));
    const uint32_t P = ctx->last_view.band_pixels;
    rk::Batch bt{};
    bt.frames = ctx->d_frames[ctx->cs].as<rk::FrameParams>() + (ctx->last_frames_offset - ctx->last_slot_frame);
    bt.rv = ctx->d_rv[ctx->last_ps].as<float4>();
    rk::k_count_visible<<<148 * 4, 256, 0, ctx->stream>>>(ctx->d_vis[ctx->last_ps].as<unsigned long long>() + (size_t)ctx->last_slot_frame * P, P, cnt);
    rk::k_count_front<<<148 * 4, 256, 0, ctx->stream>>>(ctx->scene, bt, ctx->last_slot_frame, cnt + 1);
    ctx->launches += 2;
    unsigned long long host[2] = {0, 0};
    RAST_CUDA(ctx, cudaMemcpyAsync(host, cnt, 16, cudaMemcpyDeviceToHost, ctx->stream));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->triangles = ctx->scene.T;
    out->front_facing = host[1];
    out->visible_pixels = host[0];
    out->queued_chunks = ctx->h_status.as<unsigned long long>()[0];
    return RAST_OK;
}

int rast_set_keep_visibility(rast_ctx *ctx, int enabled) {
    if (!ctx) return RAST_EINVAL;
    ctx->keep_visibility = enabled != 0;
    return RAST_OK;
}

const char *rast_last_schedule(rast_ctx *ctx) {
    if (!ctx) return "";
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->front_stream);
    cudaStreamSynchronize(ctx->stream); // the counters of the call's last batch are in h_status now
    const bool bins = ctx->sched_bins_requested && ctx->h_status.as<unsigned long long>()[rk::CNT_TILE_MODE] != 0ull;
    ctx->schedule_text = std::string("setup=") + ctx->sched_setup + " raster=" + (bins ? "k_raster_tiles" : "k_raster_chunks") + " shade=" + ctx->sched_shade +
                         " batch=" + std::to_string(ctx->sched_batch);
    return ctx->schedule_text.c_str();
}

int rast_set_retained_outputs(rast_ctx *ctx, int enabled) {
    if (!ctx) return RAST_EINVAL;
    ctx->retained_outputs = enabled != 0;
    ctx->retained.valid = false;
    ctx->retained.ext.clear();
    return RAST_OK;
}

int rast_set_profiling(rast_ctx *ctx, int enabled) {
    if (!ctx) return RAST_EINVAL;
    ctx->profiling = enabled != 0;
    return RAST_OK;
}

int rast_get_pass_ms(rast_ctx *ctx, float ms[RAST_PASS_COUNT]) {
    if (!ctx || !ms) return RAST_EINVAL;
    memcpy(ms, ctx->pass_ms, sizeof ctx->pass_ms);
    return RAST_OK;
}

uint64_t rast_launch_count(const rast_ctx *ctx) { return ctx ? ctx->launches : 0; }

int rast_selftest_division(rast_ctx *ctx, uint64_t n_samples, uint64_t seed, uint64_t *mismatches) {
    if (!ctx || !mismatches) return RAST_EINVAL;
    RAST_CUDA(ctx, cudaSetDevice(ctx->device));
    RAST_CUDA(ctx, ctx->d_aux.reserve(64));
    unsigned long long *cnt = ctx->d_aux.as<unsigned long long>();
    RAST_CUDA(ctx, cudaMemsetAsync(cnt, 0, 8, ctx->stream));
    const unsigned threads = 148u * 8u * 256u;
    const unsigned long long per_thread = (n_samples / 3 + threads - 1) / threads; // three quotients per sample triple
    rk::k_selftest_division<<<148 * 8, 256, 0, ctx->stream>>>(per_thread ? per_thread : 1ull, seed, cnt);
    ctx->launches++;
    unsigned long long host = 0;
    RAST_CUDA(ctx, cudaMemcpyAsync(&host, cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    RAST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *mismatches = host;
    return RAST_OK;
}

void *rast_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

void rast_host_free(void *p) { if (p) cudaFreeHost(p); }

int rast_host_register(void *p, uint64_t bytes) {
    if (!p || !bytes) return RAST_EINVAL;
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost) return RAST_OK; // already page-locked
    cudaGetLastError();
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return RAST_OK; }
    return e == cudaSuccess ? RAST_OK : RAST_ECUDA;
}

int rast_host_unregister(void *p) {
    if (!p) return RAST_EINVAL;
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) cudaGetLastError();
    return e == cudaSuccess ? RAST_OK : RAST_ECUDA;
}

uint64_t rast_hash64(const void *data, uint64_t bytes, uint64_t seed) {
    // multiply-xorshift over 8-byte words in four independent lanes: runs at memory speed (the same function as
    // include/rast_draw_frame.hpp's Session uses for its per-call scene fingerprint)
    const unsigned char *b = static_cast<const unsigned char *>(data);
    uint64_t l[4] = {seed ^ 0x9E3779B97F4A7C15ull, seed ^ 0xC2B2AE3D27D4EB4Full, seed ^ 0x165667B19E3779F9ull, seed ^ 0x27D4EB2F165667C5ull};
    uint64_t i = 0;
    for (; i + 32 <= bytes; i += 32) {
        uint64_t w[4];
        memcpy(w, b + i, 32);
        for (int k = 0; k < 4; ++k) { l[k] = (l[k] ^ w[k]) * 0x100000001B3ull; l[k] ^= l[k] >> 29; }
    }
    uint64_t tail[4] = {0, 0, 0, 0};
    if (i < bytes) memcpy(tail, b + i, bytes - i);
    for (int k = 0; k < 4; ++k) { l[k] = (l[k] ^ tail[k]) * 0x100000001B3ull; l[k] ^= l[k] >> 29; }
    uint64_t out = bytes;
    for (int k = 0; k < 4; ++k) { out = (out ^ l[k]) * 0xFF51AFD7ED558CCDull; out ^= out >> 33; }
    return out;
}

} // extern "C"
