// Micro-benchmark: device-to-host delivery by a KERNEL storing into mapped page-locked host memory (zero copy), against the copy engine.
// Would per-row spans written by the SMs -- which a 2-D copy cannot express -- move at PCIe speed?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zero_copy_bench zero_copy_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
// one warp per (row, plane): copies `w` bytes of row y from src to dst (both pitched W), 16 bytes per lane per step
__global__ void k_copy_rows(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, size_t pitch, unsigned w, unsigned rows, unsigned x0) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const uint4 *s = reinterpret_cast<const uint4 *>(src + (size_t)warp * pitch + x0);
    uint4 *d = reinterpret_cast<uint4 *>(dst + (size_t)warp * pitch + x0);
    for (unsigned i = lane; i < w / 16; i += 32) d[i] = s[i];
}
int main() {
    const size_t W = 1920, H = 1080, frames = 64, P = W * H;
    unsigned char *d, *h, *hd;
    cudaMalloc(&d, frames * P * 7);
    cudaHostAlloc(&h, frames * P * 7, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&hd, h, 0);
    cudaMemset(d, 1, frames * P * 7);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct { const char *name; unsigned w, rows, x0; } cases[] = {{"whole planes (1920 x 1080)", 1920, 1080, 0}, {"rect 1024 x 900", 1024, 900, 448}, {"rect 640 x 900 (spans)", 640, 900, 640}, {"rect 256 x 900", 256, 900, 832}};
    for (auto &c : cases) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0, s);
            for (size_t f = 0; f < frames; ++f) {
                const unsigned char *df = d + f * P * 7; unsigned char *hf = hd + f * P * 7;
                const unsigned blocks = (c.rows * 32 + 255) / 256;
                for (int p = 0; p < 3; ++p) k_copy_rows<<<blocks, 256, 0, s>>>(df + p * P, hf + p * P, W, c.w, c.rows, c.x0);
                k_copy_rows<<<blocks, 256, 0, s>>>(df + 3 * P, hf + 3 * P, W * 4, c.w * 4, c.rows, c.x0 * 4);
            }
            cudaEventRecord(e1, s); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)frames * c.w * c.rows * 7;
            if (rep) printf("kernel stores, %-28s %8.3f ms/frame  %7.2f GB/s of payload\n", c.name, ms / frames, bytes / ms / 1e6);
        }
    }
    // the copy engine on the same rectangles for comparison
    for (auto &c : cases) {
        cudaEventRecord(e0, s);
        for (size_t f = 0; f < frames; ++f) {
            const unsigned char *df = d + f * P * 7; unsigned char *hf = h + f * P * 7;
            for (int p = 0; p < 3; ++p) cudaMemcpy2DAsync(hf + p * P + c.x0, W, df + p * P + c.x0, W, c.w, c.rows, cudaMemcpyDeviceToHost, s);
            cudaMemcpy2DAsync(hf + 3 * P + c.x0 * 4, W * 4, df + 3 * P + c.x0 * 4, W * 4, c.w * 4, c.rows, cudaMemcpyDeviceToHost, s);
        }
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("copy engine,   %-28s %8.3f ms/frame  %7.2f GB/s of payload\n", c.name, ms / frames, (double)frames * c.w * c.rows * 7 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
