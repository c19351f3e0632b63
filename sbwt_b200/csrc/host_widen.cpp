// host_widen.cpp -- see host_widen.hpp. Host-only code (g++), no CUDA calls: the pool's submit()
// runs inside a cudaLaunchHostFunc callback, where CUDA API calls are not allowed.
#include "host_widen.hpp"

#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cstring>

#include <immintrin.h>

namespace sbwt_b200 {

// ---- int32 -> int64, non-temporal; no fence here (callers fence once per task)
__attribute__((target("avx2"))) static void widen_avx2_nofence(const int32_t* src, int64_t* dst, size_t n) {
    size_t i = 0;
    // head: bring dst to a 32-byte boundary for the stream stores
    while (i < n && ((uintptr_t)(dst + i) & 31)) { _mm_stream_si64((long long*)(dst + i), (long long)src[i]); i++; }
    for (; i + 8 <= n; i += 8) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 4));
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepi32_epi64(a));
        _mm256_stream_si256((__m256i*)(dst + i + 4), _mm256_cvtepi32_epi64(b));
    }
    for (; i < n; i++) _mm_stream_si64((long long*)(dst + i), (long long)src[i]);
}

__attribute__((target("avx2"))) static void fill_m1_i64_avx2(int64_t* dst, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { _mm_stream_si64((long long*)(dst + i), -1ll); i++; }
    const __m256i m1 = _mm256_set1_epi64x(-1);
    for (; i + 4 <= n; i += 4) _mm256_stream_si256((__m256i*)(dst + i), m1);
    for (; i < n; i++) _mm_stream_si64((long long*)(dst + i), -1ll);
}

__attribute__((target("avx2"))) static void copy_i32_avx2(const int32_t* src, int32_t* dst, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { _mm_stream_si32(dst + i, src[i]); i++; }
    for (; i + 8 <= n; i += 8) _mm256_stream_si256((__m256i*)(dst + i), _mm256_loadu_si256((const __m256i*)(src + i)));
    for (; i < n; i++) _mm_stream_si32(dst + i, src[i]);
}

__attribute__((target("avx2"))) static void fill_m1_i32_avx2(int32_t* dst, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { _mm_stream_si32(dst + i, -1); i++; }
    const __m256i m1 = _mm256_set1_epi32(-1);
    for (; i + 8 <= n; i += 8) _mm256_stream_si256((__m256i*)(dst + i), m1);
    for (; i < n; i++) _mm_stream_si32(dst + i, -1);
}

// ---- a mixed group (hits and misses): per mask byte, the next popcount(byte) packed values are permuted into the lanes of
// their hits (LUT of lane -> source index) and the miss lanes are set to -1
struct ExpandLut {
    alignas(32) int32_t idx[256][8];
    uint8_t cnt[256];
    ExpandLut() {
        for (int b = 0; b < 256; b++) {
            int k = 0;
            for (int l = 0; l < 8; l++) {
                idx[b][l] = k; // (a miss lane reads some valid element; it is overwritten with -1)
                if ((b >> l) & 1) k++;
            }
            cnt[b] = (uint8_t)k;
        }
    }
};
static const ExpandLut g_lut;

// 8 results for mask byte b: needs 8 readable int32 at src (the staging buffer is padded)
__attribute__((target("avx2"))) static inline __m256i expand8_avx2(uint32_t b, const int32_t* src) {
    const __m256i v = _mm256_loadu_si256((const __m256i*)src);
    const __m256i p = _mm256_permutevar8x32_epi32(v, _mm256_load_si256((const __m256i*)g_lut.idx[b]));
    const __m256i bits = _mm256_setr_epi32(1, 2, 4, 8, 16, 32, 64, 128);
    const __m256i miss = _mm256_cmpeq_epi32(_mm256_and_si256(_mm256_set1_epi32((int)b), bits), _mm256_setzero_si256());
    return _mm256_or_si256(p, miss);
}

__attribute__((target("avx2"))) static const int32_t* expand_group_avx2(uint32_t m, const int32_t* src, int64_t* dst) {
    const bool al = ((uintptr_t)dst & 31) == 0;
    for (int q = 0; q < 4; q++) {
        const uint32_t b = (m >> (8 * q)) & 0xFFu;
        const __m256i r = expand8_avx2(b, src);
        const __m256i lo = _mm256_cvtepi32_epi64(_mm256_castsi256_si128(r)), hi = _mm256_cvtepi32_epi64(_mm256_extracti128_si256(r, 1));
        if (al) { _mm256_stream_si256((__m256i*)(dst + 8 * q), lo); _mm256_stream_si256((__m256i*)(dst + 8 * q + 4), hi); }
        else { _mm256_storeu_si256((__m256i*)(dst + 8 * q), lo); _mm256_storeu_si256((__m256i*)(dst + 8 * q + 4), hi); }
        src += g_lut.cnt[b];
    }
    return src;
}

__attribute__((target("avx2"))) static const int32_t* expand_group_avx2(uint32_t m, const int32_t* src, int32_t* dst) {
    const bool al = ((uintptr_t)dst & 31) == 0;
    for (int q = 0; q < 4; q++) {
        const uint32_t b = (m >> (8 * q)) & 0xFFu;
        const __m256i r = expand8_avx2(b, src);
        if (al) _mm256_stream_si256((__m256i*)(dst + 8 * q), r);
        else _mm256_storeu_si256((__m256i*)(dst + 8 * q), r);
        src += g_lut.cnt[b];
    }
    return src;
}

static bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}

void widen_i32_to_i64(const int32_t* src, int64_t* dst, size_t n) {
    if (have_avx2()) { widen_avx2_nofence(src, dst, n); _mm_sfence(); return; }
    for (size_t i = 0; i < n; i++) dst[i] = src[i];
}

// run primitives per destination type
static inline void put_run(const int32_t* src, int64_t* dst, size_t n) {
    if (have_avx2()) widen_avx2_nofence(src, dst, n);
    else for (size_t i = 0; i < n; i++) dst[i] = src[i];
}
static inline void put_run(const int32_t* src, int32_t* dst, size_t n) {
    if (have_avx2()) copy_i32_avx2(src, dst, n);
    else memcpy(dst, src, n * 4);
}
static inline void put_m1(int64_t* dst, size_t n) {
    if (have_avx2()) fill_m1_i64_avx2(dst, n);
    else for (size_t i = 0; i < n; i++) dst[i] = -1;
}
static inline void put_m1(int32_t* dst, size_t n) {
    if (have_avx2()) fill_m1_i32_avx2(dst, n);
    else for (size_t i = 0; i < n; i++) dst[i] = -1;
}

// Runs of all-hit / all-miss groups (whole reads found or absent: the common case) are written with the run
// primitives above; a mixed group is written value by value.
template <typename T>
static void expand_sparse_t(const uint32_t* masks, const uint32_t* block_base, const int32_t* packed, size_t n, size_t b0, size_t b1, T* dst) {
    const size_t gpb = kSparseBlockValues / 32;
    const size_t n_groups = (n + 31) / 32;
    for (size_t b = b0; b < b1; b++) {
        const int32_t* src = packed + block_base[b];
        const size_t g_end = std::min(n_groups, (b + 1) * gpb);
        size_t g = b * gpb;
        while (g < g_end) {
            const uint32_t m = masks[g];
            const size_t first = g * 32;
            if (m == 0u || m == 0xFFFFFFFFu) {
                size_t h = g + 1;
                while (h < g_end && masks[h] == m) h++;
                const size_t cnt = std::min(n, h * 32) - first;
                if (m == 0u) put_m1(dst + first, cnt);
                else { put_run(src, dst + first, cnt); src += cnt; }
                g = h;
            } else {
                const size_t cnt = std::min<size_t>(32, n - first);
                T* o = dst + first;
                if (cnt == 32 && have_avx2()) src = expand_group_avx2(m, src, o);
                else for (size_t i = 0; i < cnt; i++) o[i] = ((m >> i) & 1u) ? (T)*src++ : (T)-1;
                g++;
            }
        }
    }
    if (have_avx2()) _mm_sfence();
}

void expand_sparse_i64(const uint32_t* masks, const uint32_t* block_base, const int32_t* packed, size_t n, size_t b0, size_t b1, int64_t* dst) {
    expand_sparse_t<int64_t>(masks, block_base, packed, n, b0, b1, dst);
}
void expand_sparse_i32(const uint32_t* masks, const uint32_t* block_base, const int32_t* packed, size_t n, size_t b0, size_t b1, int32_t* dst) {
    expand_sparse_t<int32_t>(masks, block_base, packed, n, b0, b1, dst);
}

WidenPool::WidenPool(int threads, const std::vector<int>& cpus) {
    for (int t = 0; t < threads; t++) workers_.emplace_back([this] { run(); });
    if (!cpus.empty()) { // keep the pool on the NUMA node of its GPU: the pinned buffers it reads were allocated from there
        cpu_set_t set;
        CPU_ZERO(&set);
        for (int c : cpus)
            if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &set);
        for (std::thread& w : workers_) pthread_setaffinity_np(w.native_handle(), sizeof set, &set); // (best effort)
    }
}

WidenPool::~WidenPool() {
    {
        std::lock_guard<std::mutex> g(mu_);
        stop_ = true;
    }
    work_cv_.notify_all();
    for (std::thread& w : workers_) w.join();
}

void WidenPool::push(std::vector<Task>& tasks) {
    if (tasks.empty()) return;
    {
        std::lock_guard<std::mutex> g(mu_);
        for (Task& t : tasks) {
            t.ticket->pending.fetch_add(1, std::memory_order_relaxed);
            queue_.push_back(std::move(t));
        }
    }
    work_cv_.notify_all();
}

void WidenPool::submit(const int32_t* src, int64_t* dst, size_t n, WidenTicket* t) {
    if (n == 0) return;
    const size_t T = workers_.size();
    // parts of whole 64-byte destination lines, at least 64 Ki values each
    size_t parts = std::min<size_t>(T, (n + 65535) / 65536);
    if (parts < 1) parts = 1;
    const size_t per = ((n + parts - 1) / parts + 7) & ~(size_t)7;
    std::vector<Task> tasks;
    for (size_t a = 0; a < n; a += per) {
        const size_t cnt = std::min(per, n - a);
        tasks.push_back(Task{[=] { widen_i32_to_i64(src + a, dst + a, cnt); }, t});
    }
    push(tasks);
}

void WidenPool::submit_sparse(const uint32_t* masks, const uint32_t* block_base, const int32_t* packed, size_t n, void* dst, bool dst64,
                              WidenTicket* t) {
    if (n == 0) return;
    const size_t n_blocks = (n + kSparseBlockValues - 1) / kSparseBlockValues;
    // 4 parts per worker: the cost of a block depends on its hit rate
    const size_t parts = std::max<size_t>(1, std::min<size_t>(workers_.size() * 4, (n_blocks + 15) / 16));
    const size_t per = (n_blocks + parts - 1) / parts;
    std::vector<Task> tasks;
    for (size_t b = 0; b < n_blocks; b += per) {
        const size_t e = std::min(n_blocks, b + per);
        if (dst64) tasks.push_back(Task{[=] { expand_sparse_i64(masks, block_base, packed, n, b, e, (int64_t*)dst); }, t});
        else tasks.push_back(Task{[=] { expand_sparse_i32(masks, block_base, packed, n, b, e, (int32_t*)dst); }, t});
    }
    push(tasks);
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
            task = std::move(queue_.front());
            queue_.pop_front();
        }
        task.fn();
        if (task.ticket->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> g(mu_);
            done_cv_.notify_all();
        }
    }
}

} // namespace sbwt_b200
