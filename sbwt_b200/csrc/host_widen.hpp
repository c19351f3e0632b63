// host_widen.hpp -- host half of the result wire formats of sbwt_gpu_query_host.
//
// SBWT::search / streaming_search return int64 values (SBWT.hh:390, :545). For an index with
// fewer than 2^31 columns every value fits 32 bits, and the result copy is what bounds a
// host-buffer call (8 B per k-mer over PCIe against ~0.01 ns of kernel time). Two formats:
//   dense   the device writes int32, the DMA moves half the bytes into a pinned staging buffer, and a small
//           pool of host threads sign-extends them into the caller's int64 array (non-temporal stores);
//   sparse  the device also drops the misses (aux_kernels.cuh, sparse_pack_kernel): hit masks + the hits only
//           cross PCIe, and the pool rebuilds the caller's int64 (or int32) array from them.
// Either way the values delivered are the same numbers; the work overlaps the next chunk's transfers.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace sbwt_b200 {

struct WidenTicket {
    std::atomic<int> pending{0};
};

class WidenPool {
public:
    // cpus: logical CPUs the workers may run on (empty = anywhere)
    explicit WidenPool(int threads, const std::vector<int>& cpus = std::vector<int>());
    ~WidenPool();
    int threads() const { return (int)workers_.size(); }
    // dst[i] = src[i] for i < n, split over the pool; returns at once. `t->pending` reaches 0 when done.
    void submit(const int32_t* src, int64_t* dst, size_t n, WidenTicket* t);
    // sparse format -> dst (int64 if dst64 else int32), n values, split over the pool by 4096-value blocks
    void submit_sparse(const uint32_t* masks, const uint32_t* block_base, const int32_t* packed, size_t n, void* dst, bool dst64,
                       WidenTicket* t);
    void wait(WidenTicket* t);

private:
    struct Task {
        std::function<void()> fn;
        WidenTicket* ticket;
    };
    void push(std::vector<Task>& tasks);
    void run();
    std::vector<std::thread> workers_;
    std::deque<Task> queue_;
    std::mutex mu_;
    std::condition_variable work_cv_, done_cv_;
    bool stop_ = false;
};

// single-threaded kernels of the pool (AVX2 stream stores when the CPU has them)
void widen_i32_to_i64(const int32_t* src, int64_t* dst, size_t n);
// blocks [b0, b1) of 4096 values each (the last one may be short: n values in all)
void expand_sparse_i64(const uint32_t* masks, const uint32_t* block_base, const int32_t* packed, size_t n, size_t b0, size_t b1, int64_t* dst);
void expand_sparse_i32(const uint32_t* masks, const uint32_t* block_base, const int32_t* packed, size_t n, size_t b0, size_t b1, int32_t* dst);

constexpr size_t kSparseBlockValues = 4096; // = kSparseBlock of aux_kernels.cuh

} // namespace sbwt_b200
