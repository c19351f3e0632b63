// host_widen.cpp -- see host_widen.hpp. Host-only code (g++), no CUDA calls: the pool's submit()
// runs inside a cudaLaunchHostFunc callback, where CUDA API calls are not allowed.
#include "host_widen.hpp"

#include <algorithm>

#include <immintrin.h>

namespace sbwt_b200 {

__attribute__((target("avx2"))) static void widen_avx2(const int32_t* src, int64_t* dst, size_t n) {
    size_t i = 0;
    // head: bring dst to a 32-byte boundary for the stream stores
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; i++; }
    for (; i + 8 <= n; i += 8) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 4));
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepi32_epi64(a));
        _mm256_stream_si256((__m256i*)(dst + i + 4), _mm256_cvtepi32_epi64(b));
    }
    _mm_sfence();
    for (; i < n; i++) dst[i] = src[i];
}

void widen_i32_to_i64(const int32_t* src, int64_t* dst, size_t n) {
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) { widen_avx2(src, dst, n); return; }
    for (size_t i = 0; i < n; i++) dst[i] = src[i];
}

WidenPool::WidenPool(int threads) {
    for (int t = 0; t < threads; t++) workers_.emplace_back([this] { run(); });
}

WidenPool::~WidenPool() {
    {
        std::lock_guard<std::mutex> g(mu_);
        stop_ = true;
    }
    work_cv_.notify_all();
    for (std::thread& w : workers_) w.join();
}

void WidenPool::submit(const int32_t* src, int64_t* dst, size_t n, WidenTicket* t) {
    if (n == 0) return;
    const size_t T = workers_.size();
    // parts of whole 64-byte destination lines, at least 64 Ki values each
    size_t parts = std::min<size_t>(T, (n + 65535) / 65536);
    if (parts < 1) parts = 1;
    const size_t per = ((n + parts - 1) / parts + 7) & ~(size_t)7;
    {
        std::lock_guard<std::mutex> g(mu_);
        for (size_t a = 0; a < n; a += per) {
            t->pending.fetch_add(1, std::memory_order_relaxed);
            queue_.push_back(Task{src + a, dst + a, std::min(per, n - a), t});
        }
    }
    work_cv_.notify_all();
}

void WidenPool::wait(WidenTicket* t) {
    std::unique_lock<std::mutex> g(mu_);
    done_cv_.wait(g, [t] { return t->pending.load(std::memory_order_acquire) == 0; });
}

void WidenPool::run() {
    for (;;) {
        Task task;
        {
            std::unique_lock<std::mutex> g(mu_);
            work_cv_.wait(g, [this] { return stop_ || !queue_.empty(); });
            if (queue_.empty()) return; // stop_ and drained
            task = queue_.front();
            queue_.pop_front();
        }
        widen_i32_to_i64(task.src, task.dst, task.n);
        if (task.ticket->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> g(mu_);
            done_cv_.notify_all();
        }
    }
}

} // namespace sbwt_b200
