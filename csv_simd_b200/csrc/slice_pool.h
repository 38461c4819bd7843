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
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

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

// memcpy with non-temporal stores: the staging copies of the end-to-end path move gigabytes that nobody reads back from
// cache (the DMA engine or the caller, much later), and they are bound by host memory traffic -- a plain store first
// reads the destination line (read-for-ownership), i.e. 3 units of traffic per byte copied instead of 2.  glibc switches
// to streaming stores only above a per-call threshold that the 2-4 MiB slices here stay under.
inline void stream_memcpy(void* dst, const void* src, size_t bytes)
{
#if defined(__SSE2__)
    uint8_t* d = static_cast<uint8_t*>(dst);
    const uint8_t* s = static_cast<const uint8_t*>(src);
    if (bytes < 4096) {
        std::memcpy(d, s, bytes);
        return;
    }
    const size_t head = (size_t)(-(uintptr_t)d & 63u);
    std::memcpy(d, s, head);
    d += head;
    s += head;
    bytes -= head;
    const size_t body = bytes & ~size_t(63);
    for (size_t i = 0; i < body; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
    }
    _mm_sfence();
    std::memcpy(d + body, s + body, bytes - body);
#else
    std::memcpy(dst, src, bytes);
#endif
}

// CSVB200_STREAM_COPY (A/B): bit 0 = copies INTO pinned staging stream, bit 1 = copies out of the bounce buffers into
// the caller's array stream; default 3.  Measured on two 16-vCPU B200 hosts, 1 GiB of cfg2, ms per call
// (tools/pageable_probe.py; pinned both sides: 25.6 / 26.9):
//                               mask 0         1         3
//   pageable in,  pinned out    37.5 / 38.8    35.8      32.2 / 35.8
//   pinned in,  pageable out    34.2 / 43.5    44.8      39.4 / 43.4
//   pageable both               51.1 / 60.4    55.9      41.4 / 52.0
// (the caller the path is for -- an mmap in, a Vec<usize> out -- is the last row).
inline int stream_copy_mask()
{
    static const int mask = [] {
        const char* e = std::getenv("CSVB200_STREAM_COPY");
        return e ? std::atoi(e) : 3;
    }();
    return mask;
}

enum class CopyDir { ToStaging, ToCaller };

inline void parallel_memcpy(SlicePool& pool, void* dst, const void* src, size_t bytes, CopyDir dir = CopyDir::ToCaller)
{
    const bool nt = (stream_copy_mask() & (dir == CopyDir::ToStaging ? 1 : 2)) != 0;
    static const size_t min_parallel = [] {
        const char* e = std::getenv("CSVB200_PARALLEL_COPY_MIN");   // bytes from which a copy is sliced over the pool (A/B)
        const long v = e ? std::atol(e) : 0;
        return v > 0 ? (size_t)v : (size_t)(1u << 20);
    }();
    if (bytes < min_parallel) {
        if (nt)
            stream_memcpy(dst, src, bytes);
        else
            std::memcpy(dst, src, bytes);
        return;
    }
    pool.run([&](int i, int n) {
        const size_t per = ((bytes + n - 1) / n + 63) & ~size_t(63);
        const size_t a = std::min(bytes, per * i), b = std::min(bytes, per * (i + 1));
        if (b <= a) return;
        if (nt)
            stream_memcpy(static_cast<uint8_t*>(dst) + a, static_cast<const uint8_t*>(src) + a, b - a);
        else
            std::memcpy(static_cast<uint8_t*>(dst) + a, static_cast<const uint8_t*>(src) + a, b - a);
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
