/*
 * rast.h -- C ABI of the B200-native frame path (librast_b200.so).
 *
 * Drop-in boundary for the reference renderer's frame path:
 *     void draw_frame(model_vertices, faces, model_vertnormals, vertuvs, lights, materials,
 *                     arguments, frame_buffer, depth_buffer)        headers/drawing.h:16-18
 * called from renderer.cpp:89 (single frame) and renderer.cpp:111 (spin loop).  The reference has
 * no FFI layer of its own (it is one C++ executable), so this header declares what a binding for
 * that call needs: plain pointers and sizes, no C++/torch types.  Every entry point cites the
 * reference interface it replaces.  include/rast_draw_frame.hpp holds the C++ shim with the
 * reference's own draw_frame signature on top of these calls; INTEGRATION.md shows the two-line
 * change a maintainer makes in renderer.cpp.
 *
 * Conventions: return 0 on success, a negative RAST_E* code on failure (never throws across the
 * ABI); rast_last_error() gives the text.  One context per GPU; calls on one context are
 * serialised by the caller.  There is NO CPU fallback: without a CUDA device rast_create fails.
 * Matrices are column-major 4x4 float like glm (m[c*4+r]).  Images use CImg's planar layout
 * (CImg.h:11715-11721): frame[c*W*H + y*W + x] (c = R,G,B), depth[y*W + x].
 */
#ifndef RAST_H
#define RAST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAST_OK 0
#define RAST_EINVAL (-1)   /* bad argument */
#define RAST_ECUDA (-2)    /* CUDA runtime error (no device, out of memory, launch failure) */
#define RAST_ESTATE (-3)   /* call out of order (e.g. draw before upload) */
#define RAST_ENOMEM (-4)   /* host allocation failed */

#define RAST_NO_TRIANGLE 0xFFFFFFFFu
/* rast_args.flat: 0 / 1 = the reference's -f switch, which its frame path never reads (arguments.cpp:45) -- both
 * give smooth shading, exactly like the reference.  2 = EXTENSION, not reference behaviour: one normal per face,
 * normalize(cross(c1-c0, c2-c0)) of the camera-space vertices (what readme.md:46 promises for -f; the reference
 * computes those vertices "for later use in shading", drawing.cpp:231-233, and stops there). */
#define RAST_FLAT_FACE 2

typedef struct rast_ctx rast_ctx;

/* Light -- headers/light.h:7-14.  trans_dir is an output: rast_draw_* writes
 * normalize(view * (direction,0)) into it, as Light::transform does (geometry.cpp:124-127). */
typedef struct {
    float direction[3];
    float intensity;
    float colour[3];
    float trans_dir[3];
} rast_light;

/* Material -- headers/material.h:11-25.  texels: planar f32 [3][tex_h][tex_w], already
 * normalised by normalize(0,1) as the Material constructor does (material.h:22); NULL when
 * has_texture == 0.  Copied to the device by rast_upload_materials. */
typedef struct {
    float kd[3];
    int32_t has_texture; /* 0: Kd is the albedo; 1: the texture is (the reference: material.cpp:19-21 ignores Kd of a textured material);
                          * 1 | RAST_TEXTURE_MODULATE_KD: extension, texel x Kd per channel */
    int32_t tex_w, tex_h;
    const float *texels;
} rast_material;
#define RAST_TEXTURE_MODULATE_KD 2

/* The fields of Args that draw_frame consumes -- headers/arguments.h:7-21; drawing.cpp:222
 * (scale, displacement, tait_bryan_angles), :229 (aspect_ratio), :247,255 (image size),
 * :256 (wind_clockwise).  `flat` is carried because the CLI parses it (arguments.cpp:23,45) but,
 * exactly like the reference, the path never reads it. */
typedef struct {
    uint32_t image_width, image_height;
    float aspect_ratio;
    float scale;
    float displacement[3];
    float tait_bryan_angles[3]; /* rx, ry, rz */
    int32_t wind_clockwise;
    int32_t flat;
} rast_args;

/* Per-pass device time of the last profiled frame batch, milliseconds (CUDA events on the
 * context's stream).  Filled only while rast_set_profiling(ctx, 1). */
enum { RAST_PASS_CLEAR = 0, RAST_PASS_VERTEX, RAST_PASS_SETUP, RAST_PASS_RASTER, RAST_PASS_SHADE, RAST_PASS_COUNT };

/* Workload counters of the last frame (device-side tallies, optional). */
typedef struct {
    uint64_t triangles;      /* submitted */
    uint64_t front_facing;   /* survived the cull (drawing.cpp:176-180) */
    uint64_t queued_chunks;  /* work items handed to the chunk rasteriser */
    uint64_t visible_pixels; /* pixels with a winning triangle */
} rast_stats;

/* ---- lifetime ---------------------------------------------------------------------------- */
int rast_create(int device, rast_ctx **out);
void rast_destroy(rast_ctx *ctx);
const char *rast_last_error(const rast_ctx *ctx); /* ctx may be NULL: error of the last failed rast_create */
const char *rast_version(void);

/* Launch on this cudaStream_t (passed as void*; NULL is the legacy default stream) instead of the
 * context's own non-blocking stream.  Lets a host that owns streams (e.g.
 * torch.cuda.current_stream().cuda_stream) time and order the kernels itself.
 * rast_use_own_stream switches back.
 * The passes that come before shading (parameter upload, vertex, setup, raster) run on an internal high-priority
 * stream so that they overlap the previous batch's / call's shade pass; the shade pass, the only one that writes the
 * caller's buffers, runs on this stream behind an event, so everything the caller observes is ordered on this stream.
 * RAST_OVERLAP=0 in the environment keeps every pass on this stream. */
int rast_set_stream(rast_ctx *ctx, void *cuda_stream);
int rast_use_own_stream(rast_ctx *ctx);

/* ---- scene upload (once; replaces handing the std::vectors to draw_frame each call) -------- */
/* positions xyz[n_positions], normals xyz[n_normals], uvs uv[n_uvs]; tris = 10 x int32 per
 * triangle in the order of struct Triangle (headers/face.h:6-13): v0 v1 v2 / n0 n1 n2 /
 * t0 t1 t2 / material, -1 = absent (uv, normal, material). */
int rast_upload_mesh(rast_ctx *ctx, const float *positions, uint32_t n_positions,
                     const float *normals, uint32_t n_normals, const float *uvs, uint32_t n_uvs,
                     const int32_t *tris, uint64_t n_tris);
int rast_upload_materials(rast_ctx *ctx, const rast_material *materials, uint32_t n_materials);
/* lights are kept by pointer-free copy; direction/intensity/colour are read here */
int rast_set_lights(rast_ctx *ctx, const rast_light *lights, uint32_t n_lights);

/* ---- host math (drawing.cpp:222-229, geometry.cpp:22-33,101,124-133) ----------------------- */
void rast_frame_matrices(const rast_args *args, float modelview[16], float camera[16], float normal_matrix[16], float view[16]);
void rast_transform_lights(const float view[16], rast_light *lights, uint32_t n_lights);
/* deterministic spin schedule replacing the wall-clock rotation of renderer.cpp:118-124:
 * ry_k = ry0 + (float)k * (6.2831853f / (float)n_frames) */
float rast_spin_angle(float ry0, uint32_t k, uint32_t n_frames);

/* ---- sort-first band (multi-GPU single frame) --------------------------------------------- */
/* Restrict rendering to image rows [y0, y1).  Output buffers then hold only those rows:
 * frame [3][y1-y0][W], depth [y1-y0][W].  y0 = y1 = 0 restores the whole frame. */
int rast_set_band(rast_ctx *ctx, uint32_t y0, uint32_t y1);
/* Device-pointer draws (rast_draw_frames with device_ptrs != 0) normally write planes that are exactly one band
 * large.  With a plane stride of `pixels` (>= the band) the R, G, B planes -- and consecutive frames' depth planes --
 * are `pixels` apart instead, so a band can be written straight into its rows of a full-size image: pass
 * frames = image + y0 * W, depths = depth_image + y0 * W and pixels = W * H.  0 restores the default. */
int rast_set_output_plane_stride(rast_ctx *ctx, uint64_t pixels);
/* Device-pointer draws of n frames normally fill n consecutive frame slots ([n][3][H][W], [n][H][W]).  With a frame
 * stride of `frames` the i-th frame of a call goes to slot i * frames instead: N ranks that render frame k of a sequence
 * on rank k mod N (the spin loop, renderer.cpp:105-111, partitioned by frame) each pass sequence + rank * slot_size and
 * a stride of N, and together fill ONE sequence buffer in order -- in rank 0's memory over NVLink when the pointer came
 * from rast_ipc_open.  1 restores the default. */
int rast_set_output_frame_stride(rast_ctx *ctx, uint32_t frames);

/* ---- peer memory: one process per GPU, bands / frames written straight into rank 0's image over NVLink ---------
 * The reference has no counterpart (single process, single thread).  rast_device_alloc returns plain cudaMalloc
 * memory of the context's device; rast_ipc_export turns it into a handle another process of the same node opens with
 * rast_ipc_open (cudaIpcOpenMemHandle with peer access).  The opened pointer is valid as `frames` / `depths` of a
 * device-pointer draw: the shade pass then stores its pixels into the owner's memory -- the gather IS the kernel's
 * store, no staging buffer and no collective.  The owner must not read the image before every writer's stream has
 * finished (any barrier does). */
#define RAST_IPC_HANDLE_BYTES 64
void *rast_device_alloc(rast_ctx *ctx, uint64_t bytes);
int rast_device_free(rast_ctx *ctx, void *device_ptr);
int rast_device_read(rast_ctx *ctx, void *host_dst, const void *device_src, uint64_t bytes);
int rast_ipc_export(rast_ctx *ctx, void *device_ptr, unsigned char handle[RAST_IPC_HANDLE_BYTES]);
int rast_ipc_open(rast_ctx *ctx, const unsigned char handle[RAST_IPC_HANDLE_BYTES], void **device_ptr);
int rast_ipc_close(rast_ctx *ctx, void *device_ptr);

/* ---- draw_frame (drawing.cpp:205-258) ------------------------------------------------------ */
/* Clears (frame 0, depth 1.0f: renderer.cpp:85-86,107-108), draws, and copies the result into HOST
 * buffers (pinned memory from rast_host_alloc makes the copy asynchronous and faster).  depth may
 * be NULL.  If lights_out != NULL the n_lights trans_dir values are written back
 * (geometry.cpp:126).  Returns after the images are complete in host memory. */
int rast_draw_frame(rast_ctx *ctx, const rast_args *args, uint8_t *frame, float *depth, rast_light *lights_out);

/* Same, but the result stays in DEVICE memory: frame_dev / depth_dev are device pointers (either
 * may be NULL to keep the result only in the context's internal buffers).  Asynchronous on the
 * context's stream; rast_sync waits. */
int rast_draw_frame_device(rast_ctx *ctx, const rast_args *args, uint8_t *frame_dev, float *depth_dev);

/* n frames of one scene in one call (the spin sequence): args[i] may differ in everything except
 * image size.  Frames are batched through the kernels together.  Outputs are contiguous:
 * frames [n][3][H][W], depths [n][H][W] (depths may be NULL).  `device_ptrs` != 0 means the
 * outputs are device pointers (asynchronous), else host pointers (returns when complete). */
int rast_draw_frames(rast_ctx *ctx, const rast_args *args, uint32_t n, uint8_t *frames, float *depths, int device_ptrs);

int rast_sync(rast_ctx *ctx);

/* Retained outputs (off by default).  The reference's spin loop redraws into the SAME frame / depth buffers every frame
 * (renderer.cpp:105-111).  With this switch on the caller promises that the host buffers handed to rast_draw_frame(s) are the
 * ones this context's previous host-buffer draw wrote (same pointers, same image size and band, at most as many frames) and
 * that OUTSIDE the rectangle that draw covered they still hold the cleared values -- true when they were left alone, and also
 * when the caller cleared them in between as renderer.cpp:107-108 does.  The library then rewrites only what changes: the newly covered rectangle of each frame is copied from
 * the device and the part of the previously covered rectangle outside it is reset to the cleared values (0 / 1.0f), instead of
 * rewriting the whole background of every frame on every call.  The buffers end up byte-identical to a draw without the
 * promise; a call with other buffers, another size, or the first call after switching it on is a normal full draw.  Breaking
 * the promise (something else written outside that rectangle in between) leaves it in the background. */
int rast_set_retained_outputs(rast_ctx *ctx, int enabled);

/* ---- auxiliary outputs -------------------------------------------------------------------- */
/* Winning triangle index per pixel of the most recent frame (RAST_NO_TRIANGLE = background),
 * [H][W] u32, host pointer.  This is the visibility buffer's low word; parity tests compare it
 * bit-exactly with the oracle. */
int rast_read_triangle_ids(rast_ctx *ctx, uint32_t *tri_ids);
/* The shade pass normally hands every visibility key back as "empty" while it consumes it, so that the next call needs no
 * clear pass.  With keep_visibility the LAST frame of each call keeps its keys for rast_read_triangle_ids / rast_get_stats
 * (the two return RAST_ESTATE otherwise), at the price of one clear of that frame's keys in the next call (8 bytes per
 * pixel).  Off by default: the reference has no such output. */
int rast_set_keep_visibility(rast_ctx *ctx, int enabled);
/* depth_buffer.normalize(0,255) + uchar truncation (renderer.cpp:93; CImg.h:26786-26794,52410) of
 * the most recent frame's depth, computed on the device; out = host [H][W] u8. */
int rast_depth_to_u8(rast_ctx *ctx, uint8_t *out);
int rast_get_stats(rast_ctx *ctx, rast_stats *out);
int rast_set_profiling(rast_ctx *ctx, int enabled);
int rast_get_pass_ms(rast_ctx *ctx, float ms[RAST_PASS_COUNT]);
/* Which kernel flavour each pass of the most recent batch took, e.g. "setup=k_setup<0,2> raster=k_raster_tiles
 * shade=k_resolve_shade_wt batch=240" (the library picks per batch: chunk queue or screen-tile bins, one warp per 4 rows or per tile;
 * batch = frames per launch sequence of the most recent call).
 * Waits for the context's streams; the string lives until the next call of this function on the context. */
const char *rast_last_schedule(rast_ctx *ctx);
/* number of kernel launches issued by this context since creation */
uint64_t rast_launch_count(const rast_ctx *ctx);
/* Bytes of frame / depth data this context has moved device -> host so far.  Host-buffer draws move only what holds drawn
 * pixels -- the covered span of every row, stored by a kernel straight into buffers that are page-locked and mapped
 * (rast_host_alloc, rast_host_register; image width a multiple of 16), else one rectangle per strip of rows through the copy
 * engine -- and write the constant rest of the caller's buffers (frame 0, depth 1.0f: the clear of renderer.cpp:85-86) on the
 * host, so this can be well below 7 bytes per pixel and frame; the buffers the caller sees are the same bytes either way.
 * Frames below 4 MB are copied whole.  Environment: RAST_SPARSE_COPY=0 whole frames always, RAST_DELIVER=0 never the kernel,
 * RAST_HOST_THREADS the threads that write the background (default min(4, hardware / 2)). */
uint64_t rast_d2h_bytes(rast_ctx *ctx);
/* Diagnostic: the kernels divide several numerators by one divisor with a shared reciprocal (the compiler's
 * own div.rn.f32 fast-path sequence, csrc/exact.cuh).  This compares that against IEEE division on the GPU
 * for about n_samples pseudo-random quotients and returns how many differ in any bit (expected: 0). */
int rast_selftest_division(rast_ctx *ctx, uint64_t n_samples, uint64_t seed, uint64_t *mismatches);

/* FNV-1a-64 of a host buffer: the checksum the golden fixtures of the parity tests record for frames, depth planes and
 * triangle-id maps (tests/golden/): lets any host verify a frame against them without the test infrastructure. */
uint64_t rast_fnv1a64(const void *data, uint64_t bytes);

/* pinned host memory for frame / depth buffers */
void *rast_host_alloc(uint64_t bytes);
void rast_host_free(void *p);
/* Page-lock a buffer the caller already owns (the reference's CImg frame / depth buffers live for the whole run): host-buffer
 * draws into pageable memory go through the driver's staging copies and are several times slower (a 640x480 frame: 0.49 ms per
 * call pageable).  The buffer must be unregistered before it is freed.  Registering twice is not an error. */
int rast_host_register(void *p, uint64_t bytes);
int rast_host_unregister(void *p);
/* 64-bit content hash at memory speed (not cryptographic): what the drop-in shims use to see that the scene arrays a caller
 * passes on every draw_frame call are still the ones on the device. */
uint64_t rast_hash64(const void *data, uint64_t bytes, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* RAST_H */
