// Micro-benchmark (host only): how fast can T threads write a constant background into frame / depth buffers with non-temporal stores?
// The host-side half of a host-buffer draw: g++ -O2 -pthread -o host_fill_bench host_fill_bench.cpp
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <emmintrin.h>
#include <thread>
#include <vector>
static void fill(unsigned char *b, size_t n, int v) {
    const __m128i x = _mm_set1_epi32(v);
    for (size_t i = 0; i + 64 <= n; i += 64) {
        _mm_stream_si128((__m128i *)(b + i), x); _mm_stream_si128((__m128i *)(b + i + 16), x);
        _mm_stream_si128((__m128i *)(b + i + 32), x); _mm_stream_si128((__m128i *)(b + i + 48), x);
    }
    _mm_sfence();
}
int main() {
    const size_t bytes = (size_t)512 << 20;
    unsigned char *buf = (unsigned char *)aligned_alloc(4096, bytes); if (!buf) { puts("alloc failed"); return 1; }
    memset(buf, 1, bytes);
    for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
        if (T > (int)std::thread::hardware_concurrency()) break;
        double best = 0;
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int k = 0; k < T; ++k) { const size_t per = (bytes / T) & ~(size_t)63; th.emplace_back([=]() { fill(buf + per * k, per, 0x3F800000); }); }
            for (auto &t : th) t.join();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            best = std::max(best, bytes / s / 1e9);
        }
        printf("%2d threads: %6.1f GB/s of streaming stores\n", T, best);
    }
    return 0;
}
