// Host worker pool of libcama_b200 (the loops of cama_overlay_apply_host): dynamic chunks, workers that sleep.
//
// Why not OpenMP: libgomp's workers spin for milliseconds after a parallel region, and a process only throttles that
// when IT manages more threads than there are CPUs.  With one rank per GPU on a box (torchrun) every rank's team fits
// the CPU count on its own, all of them spin, the box is oversubscribed several times over and a statically
// scheduled region waits for whichever thread lost its core: the host draw went from 0.4 ms at one rank to 11-20 ms
// at 2-8 ranks.  Here idle workers block on a condition variable (after a short spin, so that back-to-back regions
// stay hot), work is claimed chunk by chunk from an atomic counter, and a region is complete when every CHUNK is done
// — a worker that wakes up late finds nothing to do and is not waited for.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#define CAMA_CPU_RELAX() _mm_pause()
#else
#define CAMA_CPU_RELAX() std::this_thread::yield()
#endif

namespace cama {

class HostPool {
public:
    using RangeFn = void (*)(void *ctx, int64_t lo, int64_t hi);

    static HostPool &instance() {
        static HostPool *pool = new HostPool();            // never destroyed: the workers are detached and die with the process
        return *pool;
    }

    // fn(ctx, lo, hi) over [0, n) in chunks, on `threads` threads including the caller.  One region at a time.
    void run(int threads, int64_t n, int64_t chunk, RangeFn fn, void *ctx) {
        if (n <= 0) return;
        if (threads <= 1 || n <= chunk) {
            fn(ctx, 0, n);
            return;
        }
        std::lock_guard<std::mutex> one_region(run_mutex_);
        auto job = std::make_shared<Job>();
        job->fn = fn; job->ctx = ctx; job->n = n; job->chunk = chunk;
        job->total_chunks = (n + chunk - 1) / chunk;
        {
            std::lock_guard<std::mutex> lk(m_);
            while ((int)spawned_ < threads - 1) {
                std::thread(&HostPool::worker, this, (int)spawned_).detach();
                ++spawned_;
            }
            job_ = job;
            participants_ = threads - 1;
            generation_.fetch_add(1, std::memory_order_release);
        }
        cv_work_.notify_all();
        drain(*job);
        // the caller's own chunks are done; the few still in flight on other threads are short
        for (int spins = 0; job->done.load(std::memory_order_acquire) < job->total_chunks; ++spins) {
            if (spins < 4096) CAMA_CPU_RELAX();
            else std::this_thread::yield();
        }
        std::lock_guard<std::mutex> lk(m_);
        job_.reset();
    }

private:
    struct Job {
        RangeFn fn = nullptr;
        void *ctx = nullptr;
        int64_t n = 0, chunk = 1, total_chunks = 0;
        std::atomic<int64_t> next{0}, done{0};
    };

    static void drain(Job &job) {
        for (;;) {
            const int64_t c = job.next.fetch_add(1, std::memory_order_relaxed);
            if (c >= job.total_chunks) return;
            const int64_t lo = c * job.chunk, hi = lo + job.chunk < job.n ? lo + job.chunk : job.n;
            job.fn(job.ctx, lo, hi);
            job.done.fetch_add(1, std::memory_order_release);
        }
    }

    void worker(int index) {
        uint64_t seen = 0;
        for (;;) {
            for (int spins = 0; spins < 20000 && generation_.load(std::memory_order_acquire) == seen; ++spins) CAMA_CPU_RELAX();   // ~50 us
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return generation_.load(std::memory_order_acquire) != seen; });
                seen = generation_.load(std::memory_order_acquire);
                if (index < participants_) job = job_;
            }
            if (job) drain(*job);
        }
    }

    std::mutex run_mutex_, m_;
    std::condition_variable cv_work_;
    std::shared_ptr<Job> job_;
    std::atomic<uint64_t> generation_{0};
    int participants_ = 0;
    size_t spawned_ = 0;
};

}  // namespace cama
