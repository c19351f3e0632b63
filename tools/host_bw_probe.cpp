// host_bw_probe.cpp -- how fast can T host threads widen int32 -> int64 (the host side of a
// 32-bit result wire format) and memcpy? Measurement tool only.
// g++ -O3 -march=native -pthread -o host_bw_probe host_bw_probe.cpp
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

static void widen(const int32_t* src, int64_t* dst, size_t n) {
    size_t i = 0;
#ifdef __AVX2__
    for (; i + 8 <= n; i += 8) {
        __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 4));
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepi32_epi64(a));
        _mm256_stream_si256((__m256i*)(dst + i + 4), _mm256_cvtepi32_epi64(b));
    }
#endif
    for (; i < n; i++) dst[i] = src[i];
}

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : (size_t)400'000'000;
    const int only = argc > 2 ? atoi(argv[2]) : 0; // one thread count instead of the sweep
    int32_t* src = (int32_t*)aligned_alloc(64, n * 4);
    int64_t* dst = (int64_t*)aligned_alloc(64, n * 8);
    memset(src, 1, n * 4);
    memset(dst, 0, n * 8);
    std::vector<int> counts = {1, 2, 4, 8, 10, 12, 16, 24, 32};
    if (only > 0) counts = {only};
    for (int T : counts) {
        if (!only && T > (int)std::thread::hardware_concurrency() * 2) break;
        for (int what = 0; what < 2; what++) {
            double best = 1e30;
            for (int rep = 0; rep < 3; rep++) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < T; t++)
                    th.emplace_back([=]() {
                        size_t a = n * t / T & ~(size_t)7, b = (t == T - 1) ? n : (n * (t + 1) / T & ~(size_t)7);
                        if (what == 0) widen(src + a, dst + a, b - a);
                        else memcpy(dst + a / 2, src + a, (b - a) * 4);
                    });
                for (auto& x : th) x.join();
                double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (s < best) best = s;
            }
            printf("threads=%2d %s: %.2f G values/s, %.1f GB/s traffic\n", T, what == 0 ? "widen i32->i64 (stream stores)" : "memcpy i32",
                   n / best / 1e9, (what == 0 ? 12.0 : 8.0) * n / best / 1e9);
        }
    }
    return 0;
}
