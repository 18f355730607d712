// hostcopy.cpp -- thread pool + streaming copy behind the pageable-buffer path (see hostcopy.h).
#include "hostcopy.h"

#include <emmintrin.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace x266 {

namespace {

constexpr size_t SLICE = (size_t)256 << 10;    // unit of work handed to one thread

std::atomic<int> g_nt{3};
std::atomic<int> g_threadsWanted{0};

// 64 bytes per iteration, non-temporal stores: the destination is either a pinned staging slot that only the DMA engine
// reads next, or caller memory far larger than the caches -- in both cases a read-for-ownership of the line is wasted traffic.
void stream_copy(char* d, const char* s, size_t n, bool toCaller)
{
    if (!(g_nt.load(std::memory_order_relaxed) & (toCaller ? 2 : 1)) || n < 4096) { memcpy(d, s, n); return; }
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
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
    if (i < n) memcpy(d + i, s + i, n - i);
}

struct Slice {
    char* d;
    const char* s;
    size_t n;
    bool toCaller;
    std::atomic<int>* pending;
};

class Pool {
public:
    ~Pool() { shutdown(); }

    void run(const CopyJob* jobs, int nJobs)
    {
        std::atomic<int> pending{0};
        int total = 0;
        for (int j = 0; j < nJobs; j++) total += (int)((jobs[j].bytes + SLICE - 1) / SLICE);
        if (total == 0) return;
        const int want = threads_wanted();
        if (want <= 1 || total == 1) {
            for (int j = 0; j < nJobs; j++) stream_copy((char*)jobs[j].dst, (const char*)jobs[j].src, jobs[j].bytes, jobs[j].toCaller);
            return;
        }
        pending.store(total, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu_);
            grow(want - 1);
            for (int j = 0; j < nJobs; j++)
                for (size_t o = 0; o < jobs[j].bytes; o += SLICE)
                    q_.push_back(Slice{ (char*)jobs[j].dst + o, (const char*)jobs[j].src + o,
                                        jobs[j].bytes - o < SLICE ? jobs[j].bytes - o : SLICE, jobs[j].toCaller, &pending });
            queued_.store((int)q_.size(), std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire)) cv_.notify_all();
        // the caller works too (on anybody's slices), then waits for its own count to drain
        for (;;) {
            Slice sl;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (q_.empty()) break;
                sl = q_.front();
                q_.pop_front();
                queued_.store((int)q_.size(), std::memory_order_release);
            }
            work(sl);
        }
        for (int spin = 0; spin < 20000 && pending.load(std::memory_order_acquire); spin++) _mm_pause();
        std::unique_lock<std::mutex> lk(doneMu_);
        doneCv_.wait(lk, [&] { return pending.load(std::memory_order_acquire) == 0; });
    }

    void shutdown()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_)
            if (t.joinable()) t.join();
        workers_.clear();
        stop_ = false;
    }

    static int threads_wanted()
    {
        int n = g_threadsWanted.load(std::memory_order_relaxed);
        if (n > 0) return n;
        if (const char* e = getenv("X266_HOST_COPY_THREADS")) {
            n = atoi(e);
            if (n > 0) return n;
        }
        cpu_set_t set;
        int cpus = 0;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
        if (cpus <= 0) cpus = (int)std::thread::hardware_concurrency();
        n = cpus / 2;
        return n < 1 ? 1 : n > 8 ? 8 : n;
    }

private:
    void grow(int nWorkers)                       // mu_ held
    {
        while ((int)workers_.size() < nWorkers) workers_.emplace_back([this] { loop(); });
    }

    void work(const Slice& sl)
    {
        stream_copy(sl.d, sl.s, sl.n, sl.toCaller);
        if (sl.pending->fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> lk(doneMu_);
            doneCv_.notify_all();
        }
    }

    void loop()
    {
        for (;;) {
            Slice sl;
            // a staged call hands over a job every few hundred microseconds: poll briefly before going to sleep on the futex
            for (int spin = 0; spin < 40000 && !queued_.load(std::memory_order_acquire); spin++) _mm_pause();
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (q_.empty() && !stop_) {
                    sleepers_.fetch_add(1, std::memory_order_acq_rel);
                    cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                    sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                }
                if (q_.empty()) return;           // stop_ and nothing left
                sl = q_.front();
                q_.pop_front();
                queued_.store((int)q_.size(), std::memory_order_release);
            }
            work(sl);
        }
    }

    std::mutex mu_, doneMu_;
    std::condition_variable cv_, doneCv_;
    std::deque<Slice> q_;
    std::vector<std::thread> workers_;
    std::atomic<int> queued_{0}, sleepers_{0};
    bool stop_ = false;
};

Pool& pool()
{
    static Pool* p = new Pool();                  // leaked on purpose: worker threads must not outlive a destructed pool at exit
    return *p;
}

} // namespace

void host_copy_parallel(const CopyJob* jobs, int nJobs) { pool().run(jobs, nJobs); }
void set_host_copy_threads(int n) { g_threadsWanted.store(n < 0 ? 0 : n); }
int host_copy_threads() { return Pool::threads_wanted(); }
void set_host_copy_nt(int mask) { g_nt.store(mask & 3); }
void host_copy_shutdown() { pool().shutdown(); }

} // namespace x266
