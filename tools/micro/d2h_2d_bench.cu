// Micro-benchmark: device-to-host copy rate of sub-rectangles (cudaMemcpy2DAsync into pinned memory) against whole
// contiguous planes -- would copying only the covered rectangle of a sparse frame beat copying the frame?
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const size_t W = 1920, H = 1080, frames = 64;
    unsigned char *d, *h;
    cudaMalloc(&d, frames * W * H * 7);
    cudaMallocHost(&h, frames * W * H * 7);
    cudaMemset(d, 1, frames * W * H * 7);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct { const char *name; size_t w, hh; } cases[] = {{"whole frame (contiguous)", W, H}, {"rect 1000 x 900", 1000, 900}, {"rect 1920 x 900 (full rows)", W, 900}, {"rect 500 x 500", 500, 500}, {"rect 256 x 256", 256, 256}};
    for (auto &c : cases) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0, s);
            for (size_t f = 0; f < frames; ++f) {
                unsigned char *df = d + f * W * H * 7, *hf = h + f * W * H * 7;
                if (c.w == W && c.hh == H) {
                    cudaMemcpyAsync(hf, df, W * H * 7, cudaMemcpyDeviceToHost, s);
                } else {
                    for (int p = 0; p < 3; ++p) cudaMemcpy2DAsync(hf + p * W * H, W, df + p * W * H, W, c.w, c.hh, cudaMemcpyDeviceToHost, s);
                    cudaMemcpy2DAsync(hf + 3 * W * H, W * 4, df + 3 * W * H, W * 4, c.w * 4, c.hh, cudaMemcpyDeviceToHost, s);
                }
            }
            cudaEventRecord(e1, s); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)frames * c.w * c.hh * 7;
            if (rep) printf("%-30s %8.3f ms/frame  %7.2f GB/s of payload  -> %8.0f frames/s\n", c.name, ms / frames, bytes / ms / 1e6, frames / (ms * 1e-3));
        }
    }
    // the 1000 x 900 rectangle again with the frames alternating between two / four streams (several copy engines, one PCIe link)
    for (int ns : {2, 4}) {
        cudaStream_t st[4];
        for (int k = 0; k < ns; ++k) cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking);
        for (int rep = 0; rep < 2; ++rep) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s);
            for (int k = 0; k < ns; ++k) cudaStreamWaitEvent(st[k], e0, 0);
            for (size_t f = 0; f < frames; ++f) {
                unsigned char *df = d + f * W * H * 7, *hf = h + f * W * H * 7;
                cudaStream_t q = st[f % ns];
                for (int p = 0; p < 3; ++p) cudaMemcpy2DAsync(hf + p * W * H, W, df + p * W * H, W, 1000, 900, cudaMemcpyDeviceToHost, q);
                cudaMemcpy2DAsync(hf + 3 * W * H, W * 4, df + 3 * W * H, W * 4, 1000 * 4, 900, cudaMemcpyDeviceToHost, q);
            }
            cudaEvent_t done[4];
            for (int k = 0; k < ns; ++k) { cudaEventCreate(&done[k]); cudaEventRecord(done[k], st[k]); cudaStreamWaitEvent(s, done[k], 0); }
            cudaEventRecord(e1, s); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("rect 1000 x 900 on %d streams      %8.3f ms/frame  %7.2f GB/s of payload  -> %8.0f frames/s\n", ns, ms / frames, (double)frames * 1000 * 900 * 7 / ms / 1e6, frames / (ms * 1e-3));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
