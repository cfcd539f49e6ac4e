// Micro-benchmark: does the packed fp32 instruction of sm_100 (FFMA2 = fma.rn.f32x2) save issue slots?
// Four kernels with the same arithmetic per thread: scalar FFMA, packed FFMA2, and each mixed 1:1 with integer
// instructions (the shade / raster kernels are issue-bound on such a mix).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 unpack(u64 v) { float2 o; asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(v)); return o; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned iop(unsigned a, unsigned b) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, int iters, float m, float c, unsigned seed) {
    float a[8]; unsigned q[8];
    u64 p[4];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; q[i] = seed + i * threadIdx.x; }
    for (int i = 0; i < 4; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]);
    const u64 m2 = pack(m, m), c2 = pack(c, c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma1(a[i], m, c);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], m2, c2);
            }
            if (MODE >= 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) q[i] = iop(q[i], seed);
            }
        }
    }
    float s = 0; unsigned t = 0;
    for (int i = 0; i < 8; ++i) { s += a[i]; t ^= q[i]; }
    for (int i = 0; i < 4; ++i) { float2 f = unpack(p[i]); s += f.x + f.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}

int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000, blocks = 148 * 8;
    const char *names[4] = {"FFMA  x64/iter", "FFMA2 x32/iter (same flops)", "FFMA x64 + LOP3 x64", "FFMA2 x32 + LOP3 x64"};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 4; ++mode) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f, 12345u);
            if (mode == 1) k<1><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f, 12345u);
            if (mode == 2) k<2><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f, 12345u);
            if (mode == 3) k<3><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f, 12345u);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double flops = 2.0 * 64 * (double)iters * blocks * 256;
            if (rep) printf("%-32s %8.3f ms  %7.2f TFLOP/s fp32\n", names[mode], ms, flops / ms / 1e9);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
