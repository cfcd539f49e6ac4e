// Micro-benchmark: throughput of the visibility-buffer operation, a 64-bit atomicMin with unused result (RED.E.MIN.64),
// in the access patterns of the raster kernels.  MEASURED_PEAKS.json has no atomic figure, this calibrates one.
//   unique    each thread its own consecutive key, whole buffer once (coalesced: a warp = 256 contiguous bytes)
//   quad      lanes own 2x2 pixel quads of a 16x8 block like k_raster_chunks (rows W apart), blocks at random places
//   random    every thread a random key of the buffer
//   contended the `quad` pattern, each block position hit 50 times by different warps (depth complexity 50)
// for a 1080p x 32-frame buffer (531 MB), one 8K frame (265 MB) and an L2-resident 16 MB buffer.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

template <int MODE> __global__ void __launch_bounds__(256) k(u64 *vis, u64 n, unsigned W, unsigned reps, unsigned contention) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 nthreads = (u64)gridDim.x * blockDim.x;
    for (unsigned r = 0; r < reps; ++r) {
        u64 idx;
        if (MODE == 0) idx = (tid + (u64)r * nthreads) % n;
        else if (MODE == 2) idx = mix(tid * 0x9E3779B97F4A7C15ull + r) % n;
        else { // quad pattern: block origin from the warp id (divided by `contention` so that many warps share a block)
            const u64 key = (warp + (u64)r * (nthreads >> 5)) / contention;
            const u64 rows = n / W, bx = mix(key) % (W / 16), by = mix(key ^ 0x5555) % (rows / 8);
            const unsigned qx = (lane & 7) * 2, qy = (lane >> 3) * 2;
            idx = (by * 8 + qy) * W + bx * 16 + qx;
            const u64 z = (mix(warp * 77 + r) >> 20) << 32; // random depth: about ln(c) of c contenders actually lower the key
            atomicMin(vis + idx, z | (warp & 0xFFFFFFFFu)); atomicMin(vis + idx + 1, z | 1); atomicMin(vis + idx + W, z | 2); atomicMin(vis + idx + W + 1, z | 3);
            continue;
        }
        atomicMin(vis + idx, (mix(tid + r) & 0xFFFFFFFF00000000ull) | (tid & 0xFFFFFFFFu));
    }
}

int main() {
    const struct { const char *name; u64 n; unsigned W; } bufs[] = {{"1080p x 32 frames (531 MB)", 1920ull * 1080 * 32, 1920}, {"8K frame (265 MB)", 7680ull * 4320, 7680}, {"L2-resident (16 MB)", 2ull << 20, 1024}};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (auto &b : bufs) {
        u64 *vis; cudaMalloc(&vis, b.n * 8); cudaMemset(vis, 0xFF, b.n * 8);
        const unsigned blocks = 148 * 8 * 4, reps = 64;
        const double ops_simple = (double)blocks * 256 * reps, ops_quad = ops_simple * 4;
        for (int mode = 0; mode < 4; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemset(vis, 0xFF, b.n * 8);
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<blocks, 256>>>(vis, b.n, b.W, reps, 1);
                if (mode == 1) k<1><<<blocks, 256>>>(vis, b.n, b.W, reps, 1);
                if (mode == 2) k<2><<<blocks, 256>>>(vis, b.n, b.W, reps, 1);
                if (mode == 3) k<1><<<blocks, 256>>>(vis, b.n, b.W, reps, 50);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const char *names[4] = {"unique (coalesced)", "quad blocks, random places", "random addresses", "quad blocks, 50-way contention"};
            const double ops = (mode == 1 || mode == 3) ? ops_quad : ops_simple;
            printf("%-28s %-32s %8.3f ms  %7.1f G atomics/s  (%6.1f GB/s of keys)\n", b.name, names[mode], best, ops / best / 1e6, ops * 8 / best / 1e6);
        }
        cudaFree(vis);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
