// nlmc_hostpar.cpp -- host-side helpers of the boundary (plain C++, no CUDA): a persistent worker pool and the format
// conversions the class API needs around the device path.  The reference's API returns float64 arrays (M is float64
// +-1, NMC/nmc.py:52,89; NPT/npt.py:640-644); the device records int8, so the last step of a run() is a widening of up
// to a gigabyte on the host.  That step is pure memory traffic: it runs on all host threads with non-temporal stores
// (no read-for-ownership of the destination lines) and, in nlmc_msc_sweep_record_f64, chunk by chunk behind the
// device-to-host copies.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <pthread.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "nlmc_hostpar.h"

namespace nlmc {

namespace {

class WorkerPool {
  public:
    explicit WorkerPool(int n) : n_(n) {
        for (int t = 0; t < n_; ++t) threads_.emplace_back([this, t] { loop(t); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto &th : threads_) th.join();
    }
    int size() const { return n_; }
    // fn(part, parts) on `parts` <= size() workers; returns when all are done.  One job at a time (callers serialise).
    void run(int parts, const std::function<void(int, int)> &fn) {
        std::lock_guard<std::mutex> job(job_m_);
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            parts_ = parts;
            pending_ = parts;
            ++epoch_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> g(m_);
        done_cv_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void loop(int t) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)> *fn = nullptr;
            int parts = 0;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                fn = fn_;
                parts = parts_;
            }
            if (fn && t < parts) {
                (*fn)(t, parts);
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_, job_m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int, int)> *fn_ = nullptr;
    int parts_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

std::atomic<bool> g_forked_child{false};  // worker threads do not survive fork(): a child runs its parts serially

WorkerPool &pool() {
    static WorkerPool *p = [] {
        int n = (int)std::thread::hardware_concurrency();
        n = std::max(1, std::min(n, 32));
        pthread_atfork(nullptr, nullptr, [] { g_forked_child.store(true); });
        return new WorkerPool(n);  // lives until process exit (workers are detached from any handle's lifetime)
    }();
    return *p;
}

void widen_scalar(const int8_t *in, double *out, uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) out[i] = (double)in[i];
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void widen_avx2(const int8_t *in, double *out, uint64_t lo, uint64_t hi) {
    uint64_t i = lo;
    while (i < hi && (reinterpret_cast<uintptr_t>(out + i) & 31u)) { out[i] = (double)in[i]; ++i; }
    for (; i + 16 <= hi; i += 16) {
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(in + i));
        const __m256i w0 = _mm256_cvtepi8_epi32(b);                       // bytes 0..7
        const __m256i w1 = _mm256_cvtepi8_epi32(_mm_srli_si128(b, 8));    // bytes 8..15
        _mm256_stream_pd(out + i, _mm256_cvtepi32_pd(_mm256_castsi256_si128(w0)));
        _mm256_stream_pd(out + i + 4, _mm256_cvtepi32_pd(_mm256_extracti128_si256(w0, 1)));
        _mm256_stream_pd(out + i + 8, _mm256_cvtepi32_pd(_mm256_castsi256_si128(w1)));
        _mm256_stream_pd(out + i + 12, _mm256_cvtepi32_pd(_mm256_extracti128_si256(w1, 1)));
    }
    for (; i < hi; ++i) out[i] = (double)in[i];
    _mm_sfence();
}
#endif

void widen_range(const int8_t *in, double *out, uint64_t lo, uint64_t hi) {
#if defined(__x86_64__)
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    if (has_avx2) { widen_avx2(in, out, lo, hi); return; }
#endif
    widen_scalar(in, out, lo, hi);
}

int clamp_threads(int threads) {
    const int n = pool().size();
    return threads > 0 ? std::min(threads, n) : n;
}

}  // namespace

int host_threads() { return pool().size(); }

// Threads for work that EVERY rank of a multi-process job does at the same time (instance checks at handle creation):
// the pool's width divided by the number of ranks on this host (LOCAL_WORLD_SIZE, as torchrun sets it), so that eight
// ranks do not put 8 x 16 workers on 16 cores.  Work only one rank does (building M) uses the full pool.
int host_threads_shared() {
    static const int n = [] {
        int local = 1;
        if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) local = std::max(1, std::atoi(e));
        return std::max(1, pool().size() / local);
    }();
    return n;
}

void parallel_for(int parts, const std::function<void(int, int)> &fn) {
    parts = std::max(1, std::min(parts, pool().size()));
    if (parts == 1) { fn(0, 1); return; }
    if (g_forked_child.load()) {
        for (int t = 0; t < parts; ++t) fn(t, parts);
        return;
    }
    pool().run(parts, fn);
}

void widen_i8_f64(const int8_t *in, double *out, uint64_t count, int threads) {
    if (count == 0) return;
    int nt = clamp_threads(threads);
    if (count < (1u << 18)) nt = 1;
    parallel_for(nt, [&](int t, int parts) {
        const uint64_t per = ((count + (uint64_t)parts - 1) / (uint64_t)parts + 63) & ~63ull;
        const uint64_t lo = std::min(count, per * (uint64_t)t), hi = std::min(count, lo + per);
        if (lo < hi) widen_range(in, out, lo, hi);
    });
}

void prefault(void *buf, uint64_t bytes, int threads) {
    if (bytes == 0) return;
    int nt = clamp_threads(threads);
    if (bytes < (1u << 22)) nt = 1;
    volatile char *p = static_cast<volatile char *>(buf);
    parallel_for(nt, [&](int t, int parts) {
        const uint64_t per = (((bytes + (uint64_t)parts - 1) / (uint64_t)parts) + 4095) & ~4095ull;
        const uint64_t lo = std::min(bytes, per * (uint64_t)t), hi = std::min(bytes, lo + per);
        for (uint64_t i = lo; i < hi; i += 4096) p[i] = 0;
    });
}

}  // namespace nlmc
