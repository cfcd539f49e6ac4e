// Micro-benchmark: device-to-host copy rate of sub-rectangles (cudaMemcpy2DAsync into pinned memory) against whole
// contiguous planes -- would copying only the covered rectangle of a sparse frame beat copying the frame?
#include <chrono>
#include <cstdio>
#include <vector>
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
    // the same rectangle cut into K strips of rows, every strip a 3-D colour copy + a 2-D depth copy: (a) one API call per copy,
    // (b) all copies of the 64 frames in ONE cudaMemcpy3DBatchAsync call (CUDA 12.8+)
    for (int K : {1, 4, 8, 16}) {
        for (int mode = 0; mode < 2; ++mode) {
            std::vector<cudaMemcpy3DBatchOp> ops;
            float best_ms = 1e30f, best_api = 0;
            for (int rep = 0; rep < 3; ++rep) {
                ops.clear();
                cudaDeviceSynchronize();
                auto t0 = std::chrono::steady_clock::now();
                cudaEventRecord(e0, s);
                for (size_t f = 0; f < frames; ++f) {
                    unsigned char *df = d + f * W * H * 7, *hf = h + f * W * H * 7;
                    for (int k = 0; k < K; ++k) {
                        const size_t y0 = 90 + 900 * k / K, y1 = 90 + 900 * (k + 1) / K, x0 = 448, w = 1000 - 40 * (k % 3), hh = y1 - y0;
                        if (mode == 0) {
                            cudaMemcpy3DParms c3{};
                            c3.srcPtr = make_cudaPitchedPtr(df, W, W, H); c3.dstPtr = make_cudaPitchedPtr(hf, W, W, H);
                            c3.srcPos = c3.dstPos = make_cudaPos(x0, y0, 0); c3.extent = make_cudaExtent(w, hh, 3); c3.kind = cudaMemcpyDeviceToHost;
                            cudaMemcpy3DAsync(&c3, s);
                            cudaMemcpy2DAsync(hf + 3 * W * H + (y0 * W + x0) * 4, W * 4, df + 3 * W * H + (y0 * W + x0) * 4, W * 4, w * 4, hh, cudaMemcpyDeviceToHost, s);
                        } else {
                            cudaMemcpy3DBatchOp op{};
                            op.src.type = op.dst.type = cudaMemcpyOperandTypePointer;
                            op.src.op.ptr.ptr = df + y0 * W + x0; op.src.op.ptr.rowLength = W; op.src.op.ptr.layerHeight = H;
                            op.dst.op.ptr.ptr = hf + y0 * W + x0; op.dst.op.ptr.rowLength = W; op.dst.op.ptr.layerHeight = H;
                            op.extent = make_cudaExtent(w, hh, 3); op.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
                            ops.push_back(op);
                            op.src.op.ptr.ptr = df + 3 * W * H + (y0 * W + x0) * 4; op.src.op.ptr.rowLength = W * 4;
                            op.dst.op.ptr.ptr = hf + 3 * W * H + (y0 * W + x0) * 4; op.dst.op.ptr.rowLength = W * 4;
                            op.extent = make_cudaExtent(w * 4, hh, 1);
                            ops.push_back(op);
                        }
                    }
                }
                if (mode == 1) { size_t fail = 0; cudaError_t e = cudaMemcpy3DBatchAsync(ops.size(), ops.data(), &fail, 0, s); if (e != cudaSuccess) printf("batch: %s at %zu\n", cudaGetErrorString(e), fail); }
                const float api = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
                cudaEventRecord(e1, s); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best_ms) { best_ms = ms; best_api = api; }
            }
            double px = 0;
            for (int k = 0; k < K; ++k) px += (double)(1000 - 40 * (k % 3)) * ((90 + 900 * (k + 1) / K) - (90 + 900 * k / K));
            printf("%2d strips, %-22s %8.3f ms/frame  %7.2f GB/s of payload   host time of the calls %7.1f us/frame\n", K, mode ? "one batch call" : "one call per copy", best_ms / frames,
                   frames * px * 7 / best_ms / 1e6, best_api * 1e3 / frames);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
