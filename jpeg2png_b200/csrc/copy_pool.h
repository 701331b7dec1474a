// copy_pool.h — a few persistent host threads that copy between pageable caller memory and the
// pinned staging buffers (session.cu).
//
// Why not `#pragma omp parallel for` around memcpy: every staged chunk would be one fork/join of
// the OpenMP runtime.  Its workers spin between regions and then sleep; on a 128-thread host the
// wake-ups cost more than the copy (a full-width team made uploads 50x slower, profiles/
// r01_notes.md), and under a CPU quota the spinning itself starves the copy.  The pool below
// sleeps on a condition variable between jobs, never spins, and is sized once: J2P_COPY_THREADS if
// set, else min(8, CPUs this process may use / processes sharing the node), the calling thread
// included.  "CPUs this process may use" honours the cgroup quota (the GPU boxes show 128 hardware
// threads and grant 16), "processes sharing the node" is torchrun's LOCAL_WORLD_SIZE: eight ranks
// with eight copy threads each on a 16-CPU quota throttled one another (round 1: end-to-end
// weak-scaling efficiency 0.84 at 8 ranks, the device-resident number stayed at 0.998).
#pragma once
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace j2p {

class CopyPool {
public:
    static CopyPool &instance() {
        static CopyPool pool;
        return pool;
    }
    int threads() const { return nthreads_; }

    // dst[0..bytes) = src[0..bytes); returns when the copy is complete.  Re-entrant: concurrent
    // callers (compute() from several host threads) queue behind one another.
    void copy(void *dst, const void *src, size_t bytes) {
        if (nthreads_ <= 1 || bytes < (size_t)(2u << 20)) {
            memcpy(dst, src, bytes);
            return;
        }
        run(dst, src, bytes);
    }
    // First-touch every page of a freshly allocated buffer with all threads (the kernel zeroes a
    // page on its first write; done here, ahead of time, that cost overlaps the GPU's work
    // instead of sitting in the download).
    void touch(void *p, size_t bytes) {
        if (bytes == 0) return;
        if (nthreads_ <= 1 || bytes < (size_t)(2u << 20)) {
            touch_range((char *)p, bytes);
            return;
        }
        run(p, nullptr, bytes);
    }

private:
    static void touch_range(char *p, size_t bytes) {
        for (size_t off = 0; off < bytes; off += 4096) ((volatile char *)p)[off] = 0;
        ((volatile char *)p)[bytes - 1] = 0;
    }
    void run(void *dst, const void *src, size_t bytes) {
        std::lock_guard<std::mutex> serial(caller_mu_);
        // slices are multiples of 4 KB so that no two threads first-touch the same page
        const size_t slice = ((bytes + (size_t)nthreads_ - 1) / (size_t)nthreads_ + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> l(mu_);
            dst_ = (char *)dst;
            src_ = (const char *)src;
            bytes_ = bytes;
            slice_ = slice;
            pending_ = nthreads_ - 1;
            generation_++;
        }
        wake_.notify_all();
        run_slice(0);                                   // the caller copies the first slice itself
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [&] { return pending_ == 0; });
    }
    static int usable_cpus() {
        int n = 0;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
        if (n <= 0) n = (int)std::thread::hardware_concurrency();
        if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {          // "<quota> <period>" or "max <period>"
            char quota[32];
            long period = 0;
            if (fscanf(f, "%31s %ld", quota, &period) == 2 && period > 0 && strcmp(quota, "max") != 0) {
                const long q = (atol(quota) + period - 1) / period;
                if (q > 0 && q < n) n = (int)q;
            }
            fclose(f);
        }
        return n > 0 ? n : 1;
    }
    CopyPool() {
        const char *e = getenv("J2P_COPY_THREADS");
        int n = e ? atoi(e) : 0;
        if (n <= 0 || n > 64) {
            const char *lw = getenv("LOCAL_WORLD_SIZE");
            const int sharers = lw && atoi(lw) > 0 ? atoi(lw) : 1;
            n = usable_cpus() / sharers;
            if (n > 8) n = 8;
            if (n < 1) n = 1;
        }
        nthreads_ = n;
        for (int i = 1; i < nthreads_; i++) workers_.emplace_back([this, i] { worker(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
            generation_++;
        }
        wake_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void run_slice(int i) {
        const size_t off = (size_t)i * slice_;
        if (off >= bytes_) return;
        const size_t n = bytes_ - off < slice_ ? bytes_ - off : slice_;
        if (src_) memcpy(dst_ + off, src_ + off, n);
        else touch_range(dst_ + off, n);
    }
    void worker(int i) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(mu_);
                wake_.wait(l, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
            }
            run_slice(i);
            bool last;
            {
                std::lock_guard<std::mutex> l(mu_);
                last = --pending_ == 0;
            }
            if (last) done_.notify_one();
        }
    }

    int nthreads_ = 1;
    std::vector<std::thread> workers_;
    std::mutex caller_mu_, mu_;
    std::condition_variable wake_, done_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, slice_ = 0;
    int pending_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

}  // namespace j2p
