// host_widen.hpp -- host half of the 32-bit result wire format of sbwt_gpu_query_host.
//
// SBWT::search / streaming_search return int64 values (SBWT.hh:390, :545). For an index with
// fewer than 2^31 columns every value fits 32 bits, and the result copy is what bounds a
// host-buffer call (8 B per k-mer over PCIe against ~0.01 ns of kernel time), so the device
// writes int32, the DMA moves half the bytes into a pinned staging buffer, and a small pool of
// host threads sign-extends them into the caller's int64 array (non-temporal stores) while the
// next chunk is in flight. The values delivered are the same int64 numbers.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace sbwt_b200 {

struct WidenTicket {
    std::atomic<int> pending{0};
};

class WidenPool {
public:
    explicit WidenPool(int threads);
    ~WidenPool();
    int threads() const { return (int)workers_.size(); }
    // dst[i] = src[i] for i < n, split over the pool; returns at once. `t->pending` reaches 0 when done.
    void submit(const int32_t* src, int64_t* dst, size_t n, WidenTicket* t);
    void wait(WidenTicket* t);

private:
    struct Task {
        const int32_t* src;
        int64_t* dst;
        size_t n;
        WidenTicket* ticket;
    };
    void run();
    std::vector<std::thread> workers_;
    std::deque<Task> queue_;
    std::mutex mu_;
    std::condition_variable work_cv_, done_cv_;
    bool stop_ = false;
};

// single-threaded kernel of the pool (AVX2 stream stores when the CPU has them)
void widen_i32_to_i64(const int32_t* src, int64_t* dst, size_t n);

} // namespace sbwt_b200
