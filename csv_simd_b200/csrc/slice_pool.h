// slice_pool.h -- a few persistent host threads that split one job (a pread of a file range, a memcpy
// into or out of pinned memory) into slices.  One thread moves ~10 GB/s; PCIe 5 x16 wants ~55.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace csvb200 {

class SlicePool {
public:
    explicit SlicePool(int threads) : n_(std::max(1, threads))
    {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~SlicePool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // runs fn(slice_index, slice_count) on every thread (the caller is slice 0) and waits
    void run(const std::function<void(int, int)>& fn)
    {
        if (n_ == 1) {
            fn(0, 1);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            pending_ = n_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        fn(0, n_);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)>* fn = nullptr;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            if (fn) (*fn)(id, n_);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)>* fn_ = nullptr;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

inline void parallel_memcpy(SlicePool& pool, void* dst, const void* src, size_t bytes)
{
    if (bytes < (4u << 20)) {
        std::memcpy(dst, src, bytes);
        return;
    }
    pool.run([&](int i, int n) {
        const size_t per = ((bytes + n - 1) / n + 63) & ~size_t(63);
        const size_t a = std::min(bytes, per * i), b = std::min(bytes, per * (i + 1));
        if (b > a) std::memcpy(static_cast<uint8_t*>(dst) + a, static_cast<const uint8_t*>(src) + a, b - a);
    });
}

inline // CSVB200_IO_THREADS overrides; default: half the hardware threads, at most 16 (the other ranks of a
// one-process-per-GPU job need cores too)
int default_io_threads()
{
    if (const char* e = std::getenv("CSVB200_IO_THREADS")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 256) return v;
    }
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(16u, std::max(1u, hw / 2));
}


}  // namespace csvb200
