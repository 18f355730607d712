// ffi.cu -- the C ABI of libx266_b200.so (include/x266_b200.h): per-device context, the chunked
// host<->device pipeline behind the host-pointer entry points, and the Tier-1 / Tier-2 drop-in symbols
// of the reference golden model (src_tb/dct32.c, src_tb/satd.c).
//
// There is no CPU implementation of any kernel in this library.  If CUDA is unusable every entry
// point fails: int-returning ones with -1 (+ xGpuLastError()), the reference-signature void ones by
// printing the error and calling abort().
#define X266_B200_NO_GT32_DECL 1
#include "../../include/x266_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "hostcopy.h"
#include "chunk_claimer.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

namespace x266 {

// ------------------------------------------------------------------------------------------------
// bookkeeping
// ------------------------------------------------------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_dctVariant{X266_DCT_AUTO};
static std::atomic<size_t> g_dctChunk{0};          // blocks per pipeline chunk of the host-pointer DCT path (xGpuTune key 4); 0 = by batch size
static thread_local char t_err[512] = "";

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int fail(const char* what, cudaError_t e)
{
    snprintf(t_err, sizeof(t_err), "%s: %s", what, e == cudaSuccess ? "invalid argument" : cudaGetErrorString(e));
    return -1;
}

#define CK(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return fail(#call, e_);        \
    } while (0)

constexpr int MAX_DEV = 64;
constexpr int SLOTS = 4;            // chunks in flight per pipeline
constexpr int LAG = 2;              // a staged chunk is copied out to the caller LAG iterations after it was enqueued (LAG < SLOTS)
constexpr int MAX_PIPES = 4;        // concurrent host-pointer calls per device; further callers wait
constexpr int MAX_ARR = 4;          // host arrays per direction of one call

// One pipeline = what ONE host-pointer call needs: SLOTS streams with their device slots, the pinned staging ring of the
// pageable path and a buffer for call-wide inputs.  A call leases a pipeline for its duration, so two host threads on one
// GPU run on different streams and overlap instead of serialising on a per-device lock.
struct Pipe {
    cudaStream_t st[SLOTS] = {};
    cudaEvent_t evIn[SLOTS] = {};      // H2D of the slot's chunk done  -> its pinned input slot may be refilled
    cudaEvent_t evOut[SLOTS] = {};     // D2H of the slot's chunk done  -> its pinned output slot may be copied out
    void* dIn[SLOTS] = {};
    void* dOut[SLOTS] = {};
    size_t capIn[SLOTS] = {};
    size_t capOut[SLOTS] = {};
    void* pIn[SLOTS] = {};             // pinned staging (allocated on first pageable call only)
    void* pOut[SLOTS] = {};
    size_t pcapIn[SLOTS] = {};
    size_t pcapOut[SLOTS] = {};
    void* dAux = nullptr;              // persistent inputs shared by all chunks of one call (search planes, tiled frames)
    size_t capAux = 0;
};

struct Ctx {
    std::atomic<bool> ready{false};
    int dev = -1;
    int sms = 0;
    std::mutex mu;                     // guards idle / nPipes
    std::condition_variable cv;
    std::vector<Pipe*> idle;
    int nPipes = 0;
};

static Ctx g_ctx[MAX_DEV];
static std::mutex g_initMu;
static std::atomic<int> g_smCache[MAX_DEV];
static std::atomic<int> g_hostMode{0};              // xGpuTune key 12: pageable buffers 0 = staged ring, 1 = handed to the driver, 2 = cudaHostRegister per call
static std::atomic<int> g_checkModes{0};            // xGpuTune key 15: *Dev intra entry points validate mode[] on the device (debug)

int sm_count()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return 148;
    int n = g_smCache[dev].load(std::memory_order_relaxed);
    if (!n) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_smCache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

int resident_ctas_per_sm(const void* kernel, int blockThreads, size_t dynSmemBytes)
{
    struct Entry { const void* k; int dev; int block; int n; };
    static Entry cache[256];
    static int used = 0;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    std::lock_guard<std::mutex> lk(mu);
    for (int i = 0; i < used; i++)
        if (cache[i].k == kernel && cache[i].dev == dev && cache[i].block == blockThreads) return cache[i].n;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, blockThreads, dynSmemBytes) != cudaSuccess || n < 1) n = 1;
    if (used < 256) cache[used++] = Entry{ kernel, dev, blockThreads, n };
    return n;
}

static std::atomic<cudaMemPool_t> g_pool[MAX_DEV];
static std::mutex g_poolMu;

static cudaError_t pool_get(int dev, cudaMemPool_t* out)
{
    cudaMemPool_t pool = g_pool[dev].load(std::memory_order_acquire);
    if (!pool) {
        std::lock_guard<std::mutex> lk(g_poolMu);
        pool = g_pool[dev].load(std::memory_order_relaxed);
        if (!pool) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaError_t e;
            if ((e = cudaMemPoolCreate(&pool, &props)) != cudaSuccess) return e;
            unsigned long long keep = ~0ull;
            if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) { cudaMemPoolDestroy(pool); return e; }
            g_pool[dev].store(pool, std::memory_order_release);
        }
    }
    *out = pool;
    return cudaSuccess;
}

cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEV) return cudaErrorInvalidDevice;
    cudaMemPool_t pool;
    if ((e = pool_get(dev, &pool)) != cudaSuccess) return e;
    return cudaMallocFromPoolAsync(p, bytes, pool, st);
}

cudaError_t scratch_free(void* p, cudaStream_t st) { return cudaFreeAsync(p, st); }

static cudaError_t kernels_device_init() { return intra_device_init(); }
static void kernels_device_free() { intra_device_free(); }

// First use of a device: everything that allocates, uploads or configures happens HERE, so that the *Dev entry points
// really only enqueue (stream capture included): the scratch pool, the intra fragment table, the kernels' attributes.
static int ctx_get(Ctx** out)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) return fail("device index", cudaSuccess);
    Ctx& c = g_ctx[dev];
    if (!c.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(g_initMu);
        if (!c.ready.load(std::memory_order_relaxed)) {
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, dev));
            if (prop.major != 10) {
                snprintf(t_err, sizeof(t_err), "device %d is sm_%d%d; libx266_b200 is built for sm_100a only", dev, prop.major, prop.minor);
                return -1;
            }
            c.dev = dev;
            c.sms = prop.multiProcessorCount;
            cudaMemPool_t pool;
            CK(pool_get(dev, &pool));
            CK(kernels_device_init());
            c.ready.store(true, std::memory_order_release);
        }
    }
    *out = &c;
    return 0;
}

static void pipe_destroy(Pipe* p)
{
    for (int i = 0; i < SLOTS; i++) {
        if (p->st[i]) { cudaStreamSynchronize(p->st[i]); cudaStreamDestroy(p->st[i]); }
        if (p->evIn[i]) cudaEventDestroy(p->evIn[i]);
        if (p->evOut[i]) cudaEventDestroy(p->evOut[i]);
        if (p->dIn[i]) cudaFree(p->dIn[i]);
        if (p->dOut[i]) cudaFree(p->dOut[i]);
        if (p->pIn[i]) cudaFreeHost(p->pIn[i]);
        if (p->pOut[i]) cudaFreeHost(p->pOut[i]);
    }
    if (p->dAux) cudaFree(p->dAux);
    delete p;
}

static Pipe* pipe_create()
{
    Pipe* p = new Pipe();
    for (int i = 0; i < SLOTS; i++) {
        cudaError_t e = cudaStreamCreateWithFlags(&p->st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->evIn[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->evOut[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {                  // nothing half-built survives
            fail("pipeline streams/events", e);
            pipe_destroy(p);
            return nullptr;
        }
    }
    return p;
}

// RAII lease of one pipeline of the calling thread's current device.  Releasing synchronises the pipeline's streams, so
// no copy can still touch caller memory (or a staging slot) after the entry point has returned -- error paths included.
struct PipeLease {
    Ctx* c = nullptr;
    Pipe* p = nullptr;
    PipeLease()
    {
        if (ctx_get(&c)) { c = nullptr; return; }
        std::unique_lock<std::mutex> lk(c->mu);
        for (;;) {
            if (!c->idle.empty()) { p = c->idle.back(); c->idle.pop_back(); return; }
            if (c->nPipes < MAX_PIPES) { c->nPipes++; break; }
            c->cv.wait(lk);
        }
        lk.unlock();
        p = pipe_create();
        if (!p) {
            lk.lock();
            c->nPipes--;
            c->cv.notify_one();
        }
    }
    ~PipeLease()
    {
        if (!p) return;
        for (int s = 0; s < SLOTS; s++) cudaStreamSynchronize(p->st[s]);
        {
            std::lock_guard<std::mutex> lk(c->mu);
            c->idle.push_back(p);
        }
        c->cv.notify_one();
    }
    PipeLease(const PipeLease&) = delete;
    PipeLease& operator=(const PipeLease&) = delete;
    bool ok() const { return p != nullptr; }
};

static int ensure(void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

static int ensure_pinned(void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*p) CK(cudaFreeHost(*p));
    *p = nullptr; *cap = 0;
    CK(cudaHostAlloc(p, need, cudaHostAllocDefault));
    *cap = need;
    return 0;
}

// true if the DMA engines can address [p, p+bytes) directly (cudaHostAlloc / cudaHostRegister / managed memory)
static bool dma_addressable(const void* p, size_t bytes)
{
    if (!p || !bytes) return true;
    const char* ends[2] = { (const char*)p, (const char*)p + bytes - 1 };
    for (const char* q : ends) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return false; }
        if (a.type != cudaMemoryTypeHost && a.type != cudaMemoryTypeManaged) return false;
    }
    return true;
}

// One host array of a chunked call: `unit` bytes per unit, contiguous; h == nullptr means "not wanted" (skipped).
struct HostArr {
    void* h;
    size_t unit;
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------
// Chunked pipeline.  Units are independent, so chunk i runs H2D -> kernel -> D2H on stream i % SLOTS and the streams
// overlap the two copy directions with compute.
//   * caller memory the DMA engines can address (pinned / registered): copied directly, the loop only enqueues.
//   * pageable caller memory (what the reference's caller allocates, src/x266.cpp:505,647-649): staged through the
//     pipeline's pinned ring by the host copy pool -- in one parallel job per iteration, chunk i goes caller -> pinned
//     input slot while chunk i-LAG (whose D2H has completed) goes pinned output slot -> caller; the DMA of the chunks in
//     between runs underneath.  (xGpuTune 12 selects the two alternatives that were measured against it.)
// launch(dIn[k], dOut[k], first unit, units, stream) enqueues the kernel(s) of one chunk.
// ------------------------------------------------------------------------------------------------
template <typename Launch>
static int run_chunked_on(Pipe& p, const HostArr* ins, int nIn, const HostArr* outs, int nOut, ChunkClaimer& claimer, Launch launch)
{
    const size_t nUnits = claimer.nUnits, chunk = claimer.chunk;
    const int mode = g_hostMode.load(std::memory_order_relaxed);

    size_t offIn[MAX_ARR] = {}, offOut[MAX_ARR] = {}, totIn = 0, totOut = 0;
    bool stIn[MAX_ARR] = {}, stOut[MAX_ARR] = {}, anyStIn = false, anyStOut = false;
    std::vector<void*> registered;                           // xGpuTune(12, 2) only; unregistered after the streams have drained
    struct Unreg {
        std::vector<void*>& v;
        Pipe& p;
        ~Unreg()
        {
            if (v.empty()) return;
            for (int s = 0; s < SLOTS; s++) cudaStreamSynchronize(p.st[s]);
            for (void* q : v) cudaHostUnregister(q);
        }
    } unreg{ registered, p };
    auto classify = [&](const HostArr& a) -> bool {          // true = needs staging
        if (!a.h || dma_addressable(a.h, nUnits * a.unit) || mode == 1) return false;
        if (mode == 2 && cudaHostRegister(a.h, nUnits * a.unit, cudaHostRegisterDefault) == cudaSuccess) { registered.push_back(a.h); return false; }
        cudaGetLastError();
        return true;
    };
    for (int k = 0; k < nIn; k++) { offIn[k] = totIn; totIn += align256(chunk * ins[k].unit); stIn[k] = classify(ins[k]); anyStIn |= stIn[k]; }
    for (int k = 0; k < nOut; k++) { offOut[k] = totOut; if (outs[k].h) totOut += align256(chunk * outs[k].unit); stOut[k] = classify(outs[k]); anyStOut |= stOut[k]; }

    struct Fly { size_t u0, nu; int s; };
    Fly fly[SLOTS];                                          // staged chunks enqueued but not yet copied out to the caller, oldest first
    int nFly = 0;
    const bool track = anyStOut || claimer.shared;           // an event per slot marks "this slot's chunk is back on the host"
    for (size_t i = 0;; ) {
        size_t u0 = 0, nu = 0;
        // Shared counter: enqueueing is asynchronous, so without back-pressure the first thread to run would claim every chunk.  A pipeline
        // claims its next chunk only once the slot that chunk will use has drained -- a GPU behind a slower link drains, and so claims, less often.
        if (claimer.shared && i >= SLOTS) CK(cudaEventSynchronize(p.evOut[i % SLOTS]));
        const bool got = claimer.claim(&u0, &nu);
        if (!got && nFly == 0) break;
        const int s = (int)(i % SLOTS);
        CopyJob jobs[2 * MAX_ARR];
        int nj = 0;
        if (nFly && (nFly >= LAG || !got)) {                  // the oldest staged chunk is back in its pinned output slot (or will be: wait for it)
            const Fly f = fly[0];
            for (int q = 1; q < nFly; q++) fly[q - 1] = fly[q];
            nFly--;
            CK(cudaEventSynchronize(p.evOut[f.s]));
            for (int k = 0; k < nOut; k++)
                if (stOut[k]) jobs[nj++] = CopyJob{ (char*)outs[k].h + f.u0 * outs[k].unit, (char*)p.pOut[f.s] + offOut[k], f.nu * outs[k].unit, true };
        }
        if (got) {
            if (totIn && ensure(&p.dIn[s], &p.capIn[s], totIn)) return -1;
            if (totOut && ensure(&p.dOut[s], &p.capOut[s], totOut)) return -1;
            if (anyStIn) {
                if (ensure_pinned(&p.pIn[s], &p.pcapIn[s], totIn)) return -1;
                if (i >= SLOTS) CK(cudaEventSynchronize(p.evIn[s]));      // the slot's previous chunk has left host memory
                for (int k = 0; k < nIn; k++)
                    if (stIn[k]) jobs[nj++] = CopyJob{ (char*)p.pIn[s] + offIn[k], (const char*)ins[k].h + u0 * ins[k].unit, nu * ins[k].unit, false };
            }
            if (anyStOut && ensure_pinned(&p.pOut[s], &p.pcapOut[s], totOut)) return -1;
        }
        if (nj) host_copy_parallel(jobs, nj);
        if (!got) continue;
        void* dI[MAX_ARR] = {};
        void* dO[MAX_ARR] = {};
        for (int k = 0; k < nIn; k++) {
            dI[k] = (char*)p.dIn[s] + offIn[k];
            const void* from = stIn[k] ? (const void*)((char*)p.pIn[s] + offIn[k]) : (const void*)((const char*)ins[k].h + u0 * ins[k].unit);
            CK(cudaMemcpyAsync(dI[k], from, nu * ins[k].unit, cudaMemcpyHostToDevice, p.st[s]));
        }
        if (anyStIn) CK(cudaEventRecord(p.evIn[s], p.st[s]));
        for (int k = 0; k < nOut; k++) dO[k] = outs[k].h ? (char*)p.dOut[s] + offOut[k] : nullptr;
        CK(launch(dI, dO, u0, nu, p.st[s]));
        for (int k = 0; k < nOut; k++) {
            if (!outs[k].h) continue;
            void* to = stOut[k] ? (void*)((char*)p.pOut[s] + offOut[k]) : (void*)((char*)outs[k].h + u0 * outs[k].unit);
            CK(cudaMemcpyAsync(to, dO[k], nu * outs[k].unit, cudaMemcpyDeviceToHost, p.st[s]));
        }
        if (track) CK(cudaEventRecord(p.evOut[s], p.st[s]));
        if (anyStOut) fly[nFly++] = Fly{ u0, nu, s };         // nFly <= LAG < SLOTS here: the slot about to be reused is never in flight
        i++;
    }
    for (int s = 0; s < SLOTS; s++) CK(cudaStreamSynchronize(p.st[s]));
    return 0;
}

template <typename Launch>
static int run_chunked_on(Pipe& p, const HostArr* ins, int nIn, const HostArr* outs, int nOut, size_t nUnits, size_t unitsPerChunk, Launch launch)
{
    ChunkClaimer claimer(nUnits, nUnits < unitsPerChunk ? nUnits : unitsPerChunk);
    return run_chunked_on(p, ins, nIn, outs, nOut, claimer, launch);
}

template <typename Launch>
static int run_chunked(const HostArr* ins, int nIn, const HostArr* outs, int nOut, size_t nUnits, size_t unitsPerChunk, Launch launch)
{
    PipeLease lease;
    if (!lease.ok()) return -1;
    return run_chunked_on(*lease.p, ins, nIn, outs, nOut, nUnits, unitsPerChunk, launch);
}

// the common case: one array in, one array out
template <typename Launch>
static int run_chunked(const void* src, size_t inUnit, void* dst, size_t outUnit, size_t nUnits, size_t unitsPerChunk, Launch launch)
{
    const HostArr in{ const_cast<void*>(src), inUnit }, out{ dst, outUnit };
    return run_chunked(&in, 1, &out, 1, nUnits, unitsPerChunk,
                       [&](void* const* dI, void* const* dO, size_t, size_t n, cudaStream_t st) { return launch(dI[0], dO[0], n, st); });
}

static cudaError_t dct32_dispatch(const int16_t* s, int16_t* d, size_t n, int s1, int s2, cudaStream_t st)
{
    const int v = g_dctVariant.load(std::memory_order_relaxed);
    if (v == X266_DCT_BFLY) return launch_dct32_bfly(s, d, n, s1, s2, st);
    return launch_dct32_imma(s, d, n, s1, s2, st);
}

static bool shifts_ok(int s1, int s2) { return s1 >= 1 && s1 <= 16 && s2 >= 1 && s2 <= 16; }

[[noreturn]] static void die(const char* fn)
{
    fprintf(stderr, "libx266_b200: %s failed: %s (no CPU fallback exists)\n", fn, t_err);
    abort();
}

} // namespace x266

using namespace x266;

// ================================================================================================
// g_t32 (replaces src_tb/dct32.c:30-64).  Same layout as `const short g_t32[32][32]`.
// ================================================================================================
struct x266_g16_t { short v[32][32]; };
static constexpr x266_g16_t make_g16()
{
    x266_g16_t g{};
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) g.v[k][n] = (short)g32(k, n);
    return g;
}
extern "C" {
extern __attribute__((visibility("default"))) const x266_g16_t g_t32;
const x266_g16_t g_t32 = make_g16();
}

// ================================================================================================
// Tier 3
// ================================================================================================
extern "C" int xGpuInit(int device)
{
    if (device >= 0) CK(cudaSetDevice(device));
    Ctx* c;
    return ctx_get(&c);
}

extern "C" void xGpuFree(void)
{
    // mirrors xCodecFree (src/x266.cpp:515-524): the caller has no call in flight on this device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return;
    Ctx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lk(g_initMu);
    cudaDeviceSynchronize();
    {
        std::lock_guard<std::mutex> lp(g_poolMu);
        cudaMemPool_t pool = g_pool[dev].exchange(nullptr);
        if (pool) cudaMemPoolDestroy(pool);        // scratch of the search kernels
    }
    if (!c.ready.load(std::memory_order_acquire)) return;
    {
        std::lock_guard<std::mutex> lc(c.mu);
        for (Pipe* p : c.idle) pipe_destroy(p);
        c.nPipes -= (int)c.idle.size();
        c.idle.clear();
    }
    kernels_device_free();
    c.ready.store(false, std::memory_order_release);
}

extern "C" const char* xGpuLastError(void) { return t_err; }

extern "C" int xGpuHostRegister(void* p, size_t bytes)
{
    if (!p || !bytes) return fail("xGpuHostRegister", cudaSuccess);
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return 0;
}

extern "C" int xGpuHostCopyThreads(void) { return host_copy_threads(); }

extern "C" int xGpuHostUnregister(void* p)
{
    if (!p) return fail("xGpuHostUnregister", cudaSuccess);
    CK(cudaHostUnregister(p));
    return 0;
}

extern "C" int xIntra32MmaTable(uint32_t* table)
{
    if (!table) return fail("xIntra32MmaTable", cudaSuccess);
    intra_mma_table_copy(table);                 // pure host computation: works without a device
    return 0;
}
extern "C" unsigned long long xGpuKernelLaunches(void) { return g_launches.load(); }

extern "C" int xGpuSetDctVariant(int variant)
{
    if (variant < X266_DCT_AUTO || variant > X266_DCT_IMMA) return fail("xGpuSetDctVariant", cudaSuccess);
    g_dctVariant.store(variant);
    return 0;
}

extern "C" int xGpuTune(int key, int value)
{
    // diagnostic hook used by scripts/tune_dct.py; key 0 = IMMA kernel instantiation id
    if (key == 0) { set_imma_config(value); return 0; }
    if (key == 1) { set_search_v1(value); return 0; }
    if (key == 2) { set_satd_cuda_cores(value); return 0; }
    if (key == 3) { set_small_dct_cuda_cores(value); return 0; }
    if (key == 5) { set_decide_v1(value); return 0; }
    if (key == 6) { set_search_acc_form(value); return 0; }
    if (key == 7) { set_sad_search_v1(value); return 0; }
    if (key == 8) { set_intra_swar(value); return 0; }
    if (key == 9) { set_intra_ctas(value); return 0; }
    if (key == 10) { set_dct8_ctas(value); return 0; }
    if (key == 11) { set_dct4_ctas(value); return 0; }
    if (key == 4 && value >= 0) { g_dctChunk.store((size_t)value); return 0; }
    if (key == 12 && value >= 0 && value <= 2) { g_hostMode.store(value); return 0; }
    if (key == 13 && value >= 0) { set_host_copy_threads(value); return 0; }
    if (key == 14) { set_host_copy_nt(value); return 0; }
    if (key == 15) { g_checkModes.store(value ? 1 : 0); return 0; }
    if (key == 16) { set_frame_resi_config(value); return 0; }
    return fail("xGpuTune: unknown key", cudaSuccess);
}

// *Dev entry points: the device must have been initialised (pool, tables, attributes) before they may "only enqueue";
// the check is one atomic load after the first call.
static int dev_ready()
{
    Ctx* c;
    return ctx_get(&c);
}

extern "C" int xDct32BatchDev(const int16_t* dSrc, int16_t* dDst, size_t nBlocks, int s1, int s2, void* stream)
{
    if (!shifts_ok(s1, s2) || (nBlocks && (!dSrc || !dDst))) return fail("xDct32BatchDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dSrc) | reinterpret_cast<uintptr_t>(dDst)) & 15) return fail("xDct32BatchDev: 16-byte alignment", cudaSuccess);
    if (dev_ready()) return -1;
    CK(dct32_dispatch(dSrc, dDst, nBlocks, s1, s2, (cudaStream_t)stream));
    return 0;
}

static size_t dct32_chunk(size_t nBlocks)
{
    // 32 MiB chunks amortise the per-chunk hand-over best (47.0 GB/s each way of the 48.2 the link gives with both directions busy);
    // below ~24 chunks the fill and drain of the pipeline cost more than that, so smaller batches use 16 Ki blocks
    // (profiles/r01_e2e_chunk_sweep.log).
    const size_t chunk = g_dctChunk.load();
    return chunk ? chunk : nBlocks >= ((size_t)3 << 18) ? 32768 : 16384;
}

extern "C" int xDct32Batch(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2)
{
    if (!shifts_ok(s1, s2) || (nBlocks && (!src || !dst))) return fail("xDct32Batch", cudaSuccess);
    if (nBlocks == 0) return 0;
    return run_chunked(src, 2048, dst, 2048, nBlocks, dct32_chunk(nBlocks),
                       [&](void* di, void* dO, size_t n, cudaStream_t st) {
                           return dct32_dispatch((const int16_t*)di, (int16_t*)dO, n, s1, s2, st);
                       });
}

extern "C" int xDct32BatchMultiGpu(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, int nGpus)
{
    // blocks are independent, no collective, no peer traffic.  One host thread per device drives that device's
    // chunked pipeline; the threads claim chunks from ONE shared counter instead of owning a fixed [g*N/G, (g+1)*N/G) range,
    // because the host links of a box are not equal (on the 8-GPU pool box GPUs 4-7 get 36 GB/s each way, GPUs 0-3 47, and
    // all of them together 64: profiles/r02_link_ceiling.md) -- a static split finishes with the slowest link.
    int have = 0;
    CK(cudaGetDeviceCount(&have));
    if (nGpus <= 0) nGpus = have;
    if (nGpus > have || !shifts_ok(s1, s2) || (nBlocks && (!src || !dst))) return fail("xDct32BatchMultiGpu", cudaSuccess);
    if (nBlocks == 0) return 0;
    int prev = 0;
    CK(cudaGetDevice(&prev));
    // chunks small enough that every GPU gets several, large enough to amortise the hand-over
    size_t chunk = dct32_chunk(nBlocks / nGpus + 1);
    ChunkClaimer claimer(nBlocks, chunk < nBlocks ? chunk : nBlocks, /*shared*/ true);
    const HostArr in{ const_cast<int16_t*>(src), 2048 }, out{ dst, 2048 };
    std::vector<int> rc(nGpus, 0);
    std::vector<std::string> err(nGpus);
    std::vector<std::thread> th;
    for (int g = 0; g < nGpus; g++)
        th.emplace_back([&, g]() {
            if (cudaSetDevice(g) != cudaSuccess) { rc[g] = -1; err[g] = "cudaSetDevice failed"; return; }
            PipeLease lease;
            rc[g] = !lease.ok() ? -1
                  : run_chunked_on(*lease.p, &in, 1, &out, 1, claimer,
                                   [&](void* const* dI, void* const* dO, size_t, size_t n, cudaStream_t st) {
                                       return dct32_dispatch((const int16_t*)dI[0], (int16_t*)dO[0], n, s1, s2, st);
                                   });
            if (rc[g]) err[g] = t_err;
        });
    for (auto& t : th) t.join();
    cudaSetDevice(prev);
    for (int g = 0; g < nGpus; g++)
        if (rc[g]) { snprintf(t_err, sizeof(t_err), "xDct32BatchMultiGpu: device %d: %s", g, err[g].c_str()); return -1; }
    return 0;
}

extern "C" int xIdct32BatchDev(const int16_t* dSrc, int16_t* dDst, size_t nBlocks, int s1, int s2, void* stream)
{
    if (!shifts_ok(s1, s2) || (nBlocks && (!dSrc || !dDst))) return fail("xIdct32BatchDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dSrc) | reinterpret_cast<uintptr_t>(dDst)) & 15) return fail("xIdct32BatchDev: 16-byte alignment", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_idct32_imma(dSrc, dDst, nBlocks, s1, s2, (cudaStream_t)stream));
    return 0;
}

extern "C" int xIdct32Batch(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2)
{
    if (!shifts_ok(s1, s2) || (nBlocks && (!src || !dst))) return fail("xIdct32Batch", cudaSuccess);
    if (nBlocks == 0) return 0;
    return run_chunked(src, 2048, dst, 2048, nBlocks, 8192,
                       [&](void* di, void* dO, size_t n, cudaStream_t st) {
                           return launch_idct32_imma((const int16_t*)di, (int16_t*)dO, n, s1, s2, st);
                       });
}

extern "C" int xDctNBatchDev(int log2N, const int16_t* dSrc, int16_t* dDst, size_t nBlocks, int s1, int s2, void* stream)
{
    if (log2N == 5) return xDct32BatchDev(dSrc, dDst, nBlocks, s1, s2, stream);
    if (log2N < 2 || log2N > 5 || !shifts_ok(s1, s2) || (nBlocks && (!dSrc || !dDst))) return fail("xDctNBatchDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dSrc) | reinterpret_cast<uintptr_t>(dDst)) & 15) return fail("xDctNBatchDev: 16-byte alignment", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_dctN(log2N, dSrc, dDst, nBlocks, s1, s2, (cudaStream_t)stream));
    return 0;
}

extern "C" int xDctNBatch(int log2N, const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2)
{
    if (log2N == 5) return xDct32Batch(src, dst, nBlocks, s1, s2);
    if (log2N < 2 || log2N > 5 || !shifts_ok(s1, s2) || (nBlocks && (!src || !dst))) return fail("xDctNBatch", cudaSuccess);
    if (nBlocks == 0) return 0;
    const size_t unit = (size_t)2 << (2 * log2N);
    return run_chunked(src, unit, dst, unit, nBlocks, (size_t)(16u << 20) / unit,
                       [&](void* di, void* dO, size_t n, cudaStream_t st) {
                           return launch_dctN(log2N, (const int16_t*)di, (int16_t*)dO, n, s1, s2, st);
                       });
}

extern "C" int xPartialButterfly32Dev(const int16_t* dSrc, int16_t* dDst, int shift, int line, void* stream)
{
    if (shift < 1 || shift > 16 || line < 0 || (line && (!dSrc || !dDst))) return fail("xPartialButterfly32Dev", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_partial32(dSrc, dDst, shift, line, (cudaStream_t)stream));
    return 0;
}

extern "C" int xSatd8x8BatchDev(const int16_t* dDiff, int32_t* dSatd, size_t n, void* stream)
{
    if (n && (!dDiff || !dSatd)) return fail("xSatd8x8BatchDev", cudaSuccess);
    if (reinterpret_cast<uintptr_t>(dDiff) & 15) return fail("xSatd8x8BatchDev: 16-byte alignment", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_satd8x8_batch(dDiff, dSatd, n, (cudaStream_t)stream));
    return 0;
}

extern "C" int xSatd8x8Batch(const int16_t* diff, int32_t* satd, size_t n)
{
    if (n && (!diff || !satd)) return fail("xSatd8x8Batch", cudaSuccess);
    if (n == 0) return 0;
    return run_chunked(diff, 128, satd, 4, n, (size_t)1 << 17,
                       [&](void* di, void* dO, size_t m, cudaStream_t st) {
                           return launch_satd8x8_batch((const int16_t*)di, (int32_t*)dO, m, st);
                       });
}

// CT = element type of the cost surface: uint32_t (the SURVEY 8(d) config-3 layout) or uint16_t (the ...U16 entry points)
template <typename CT>
using search_launch_t = cudaError_t (*)(const uint8_t*, const uint8_t*, intptr_t, int, int, int, size_t, size_t, CT*, int32_t*, cudaStream_t);

template <typename CT>
static int search_dev(const char* api, search_launch_t<CT> launch, const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h,
                      int range, size_t blk0, size_t blk1, CT* dCost, int32_t* dBest, void* stream)
{
    if (!dCur || !dRefPadded || w <= 0 || h <= 0 || (w & 7) || (h & 7) || range < 0 || strd < w + 2 * range || blk1 < blk0 ||
        blk1 > (size_t)(w / 8) * (h / 8))
        return fail(api, cudaSuccess);
    if (blk1 == blk0) return 0;
    if (dev_ready()) return -1;
    cudaError_t e = launch(dCur, dRefPadded, strd, w, h, range, blk0, blk1, dCost, dBest, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(api, e);
    return 0;
}

// Host form of both searches: the two planes go up once (call-wide inputs), then the block range runs through the
// chunked pipeline with no per-chunk input and up to two outputs (cost surface, argmin triples).
template <typename CT>
static int search_host(const char* api, search_launch_t<CT> launch, const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h,
                       int range, size_t blk0, size_t blk1, CT* cost, int32_t* best)
{
    if (!cur || !refPadded || w <= 0 || h <= 0 || (w & 7) || (h & 7) || range < 0 || strd < w + 2 * range || blk1 < blk0 ||
        blk1 > (size_t)(w / 8) * (h / 8))
        return fail(api, cudaSuccess);
    if (blk1 == blk0) return 0;
    const size_t curBytes = (size_t)w * h;
    const size_t refBytes = (size_t)strd * (h + 2 * range);
    const size_t refOff = align256(curBytes);
    PipeLease l;
    if (!l.ok()) return -1;
    Pipe& p = *l.p;
    if (ensure(&p.dAux, &p.capAux, refOff + refBytes)) return -1;
    uint8_t* dCur = (uint8_t*)p.dAux;
    uint8_t* dRef = dCur + refOff;
    CK(cudaMemcpyAsync(dCur, cur, curBytes, cudaMemcpyHostToDevice, p.st[0]));
    CK(cudaMemcpyAsync(dRef, refPadded, refBytes, cudaMemcpyHostToDevice, p.st[0]));
    CK(cudaStreamSynchronize(p.st[0]));
    const size_t side = (size_t)(2 * range + 1);
    const size_t costUnit = side * side * sizeof(CT);
    // blocks per chunk: keep a chunk's cost surface around 64 MiB
    size_t per = cost ? ((size_t)64 << 20) / costUnit : (size_t)1 << 20;
    if (per < 1) per = 1;
    const HostArr outs[2] = { { cost, costUnit }, { best, 12 } };
    if (run_chunked_on(p, nullptr, 0, outs, 2, blk1 - blk0, per,
                       [&](void* const*, void* const* dO, size_t u0, size_t nu, cudaStream_t st) {
                           return launch(dCur, dRef, strd, w, h, range, blk0 + u0, blk0 + u0 + nu, (CT*)dO[0], (int32_t*)dO[1], st);
                       })) {
        const std::string inner(t_err);
        snprintf(t_err, sizeof(t_err), "%s: %s", api, inner.c_str());
        return -1;
    }
    return 0;
}

extern "C" int xSatd8x8SearchDev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                                 size_t blk0, size_t blk1, uint32_t* dCost, int32_t* dBest, void* stream)
{
    return search_dev<uint32_t>("xSatd8x8SearchDev", launch_satd8x8_search, dCur, dRefPadded, strd, w, h, range, blk0, blk1, dCost, dBest, stream);
}

extern "C" int xSatd8x8Search(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                              size_t blk0, size_t blk1, uint32_t* cost, int32_t* best)
{
    return search_host<uint32_t>("xSatd8x8Search", launch_satd8x8_search, cur, refPadded, strd, w, h, range, blk0, blk1, cost, best);
}

extern "C" int xSad8x8Search(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                             size_t blk0, size_t blk1, uint32_t* cost, int32_t* best)
{
    return search_host<uint32_t>("xSad8x8Search", launch_sad8x8_search, cur, refPadded, strd, w, h, range, blk0, blk1, cost, best);
}

extern "C" int xSad8x8SearchDev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                                size_t blk0, size_t blk1, uint32_t* dCost, int32_t* dBest, void* stream)
{
    return search_dev<uint32_t>("xSad8x8SearchDev", launch_sad8x8_search, dCur, dRefPadded, strd, w, h, range, blk0, blk1, dCost, dBest, stream);
}

// 16-bit cost surfaces: the same searches writing uint16_t costs (exact: an 8x8 SATD of 8-bit pixels is <= 32640, a SAD <= 16320) --
// half the bytes of the u32 surface that dominates the searches' HBM traffic (547.6 -> 273.8 MB per 1080p +-32 frame).
extern "C" int xSatd8x8SearchU16Dev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                                    size_t blk0, size_t blk1, uint16_t* dCost, int32_t* dBest, void* stream)
{
    return search_dev<uint16_t>("xSatd8x8SearchU16Dev", launch_satd8x8_search, dCur, dRefPadded, strd, w, h, range, blk0, blk1, dCost, dBest, stream);
}

extern "C" int xSatd8x8SearchU16(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                                 size_t blk0, size_t blk1, uint16_t* cost, int32_t* best)
{
    return search_host<uint16_t>("xSatd8x8SearchU16", launch_satd8x8_search, cur, refPadded, strd, w, h, range, blk0, blk1, cost, best);
}

extern "C" int xSad8x8SearchU16Dev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                                   size_t blk0, size_t blk1, uint16_t* dCost, int32_t* dBest, void* stream)
{
    return search_dev<uint16_t>("xSad8x8SearchU16Dev", launch_sad8x8_search, dCur, dRefPadded, strd, w, h, range, blk0, blk1, dCost, dBest, stream);
}

extern "C" int xSad8x8SearchU16(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                                size_t blk0, size_t blk1, uint16_t* cost, int32_t* best)
{
    return search_host<uint16_t>("xSad8x8SearchU16", launch_sad8x8_search, cur, refPadded, strd, w, h, range, blk0, blk1, cost, best);
}

extern "C" int sad(unsigned char* input_data1, unsigned char* input_data2, size_t n)
{
    // replaces riscv/programs/benchmarks/sad/sad.c:27-38 (host pointers, synchronous)
    const size_t bytes = n * n;
    unsigned out = 0;
    auto body = [&]() -> int {
        PipeLease l;
        if (!l.ok()) return -1;
        Pipe& p = *l.p;
        if (ensure(&p.dIn[0], &p.capIn[0], 2 * bytes + 256)) return -1;
        if (ensure(&p.dOut[0], &p.capOut[0], 256)) return -1;
        uint8_t* dA = (uint8_t*)p.dIn[0];
        uint8_t* dB = dA + align256(bytes);
        CK(cudaMemcpyAsync(dA, input_data1, bytes, cudaMemcpyHostToDevice, p.st[0]));
        CK(cudaMemcpyAsync(dB, input_data2, bytes, cudaMemcpyHostToDevice, p.st[0]));
        CK(launch_sad_region(dA, dB, bytes, (unsigned*)p.dOut[0], p.st[0]));
        CK(cudaMemcpyAsync(&out, p.dOut[0], sizeof(out), cudaMemcpyDeviceToHost, p.st[0]));
        CK(cudaStreamSynchronize(p.st[0]));
        return 0;
    };
    if (body()) die("sad");
    return (int)out;
}

// optional device-side validation of the mode array behind xGpuTune(15, 1): the *Dev intra entry points otherwise trust the
// caller (the host forms always check), and an out-of-range mode is predicted as DC
static int check_modes_dev(const char* api, const uint8_t* dMode, size_t n, cudaStream_t st)
{
    if (!g_checkModes.load(std::memory_order_relaxed) || !n) return 0;
    unsigned* dBad = nullptr;
    unsigned bad = 0;
    CK(scratch_alloc((void**)&dBad, sizeof(unsigned), st));
    CK(cudaMemsetAsync(dBad, 0, sizeof(unsigned), st));
    cudaError_t e = launch_mode_range_check(dMode, n, 34, dBad, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, dBad, sizeof(bad), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    scratch_free(dBad, st);
    if (e != cudaSuccess) return fail(api, e);
    if (bad) { snprintf(t_err, sizeof(t_err), "%s: %u mode value(s) > 34", api, bad); return -1; }
    return 0;
}

extern "C" int xIntra32PredDev(const uint8_t* dRefs, const uint8_t* dMode, uint8_t* dPred, size_t n, void* stream)
{
    if (n && (!dRefs || !dMode || !dPred)) return fail("xIntra32PredDev", cudaSuccess);
    if (reinterpret_cast<uintptr_t>(dPred) & 15) return fail("xIntra32PredDev: 16-byte alignment of pred", cudaSuccess);
    if (dev_ready()) return -1;
    if (check_modes_dev("xIntra32PredDev", dMode, n, (cudaStream_t)stream)) return -1;
    CK(launch_intra32(dRefs, dMode, dPred, n, (cudaStream_t)stream));
    return 0;
}

extern "C" int xIntra32Pred(const uint8_t* refs, const uint8_t* mode, uint8_t* pred, size_t n)
{
    if (n && (!refs || !mode || !pred)) return fail("xIntra32Pred", cudaSuccess);
    for (size_t i = 0; i < n; i++)
        if (mode[i] > 34) return fail("xIntra32Pred: mode > 34", cudaSuccess);
    if (n == 0) return 0;
    const HostArr ins[2] = { { const_cast<uint8_t*>(refs), 129 }, { const_cast<uint8_t*>(mode), 1 } };
    const HostArr out{ pred, 1024 };
    return run_chunked(ins, 2, &out, 1, n, (size_t)1 << 15,
                       [&](void* const* dI, void* const* dO, size_t, size_t np, cudaStream_t st) {
                           return launch_intra32((const uint8_t*)dI[0], (const uint8_t*)dI[1], (uint8_t*)dO[0], np, st);
                       });
}

extern "C" int xIntra32PredModesDev(const uint8_t* dRefs, size_t nBlocks, uint64_t modeMask, uint8_t* dPred, void* stream)
{
    if ((modeMask >> 35) || (nBlocks && modeMask && (!dRefs || !dPred))) return fail("xIntra32PredModesDev", cudaSuccess);
    if (reinterpret_cast<uintptr_t>(dPred) & 15) return fail("xIntra32PredModesDev: 16-byte alignment of pred", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_intra32_modes(dRefs, modeMask, dPred, nBlocks, (cudaStream_t)stream));
    return 0;
}

extern "C" int xIntra32PredModes(const uint8_t* refs, size_t nBlocks, uint64_t modeMask, uint8_t* pred)
{
    if ((modeMask >> 35) || (nBlocks && modeMask && (!refs || !pred))) return fail("xIntra32PredModes", cudaSuccess);
    const size_t nModes = (size_t)__builtin_popcountll(modeMask);
    if (nBlocks == 0 || nModes == 0) return 0;
    const HostArr in{ const_cast<uint8_t*>(refs), 129 }, out{ pred, nModes * 1024 };
    return run_chunked(&in, 1, &out, 1, nBlocks, ((size_t)1 << 15) / nModes + 1,
                       [&](void* const* dI, void* const* dO, size_t, size_t nb, cudaStream_t st) {
                           return launch_intra32_modes((const uint8_t*)dI[0], modeMask, (uint8_t*)dO[0], nb, st);
                       });
}

extern "C" int xTranspose32x32BatchDev(const uint8_t* dSrc, uint8_t* dDst, size_t nTiles, void* stream)
{
    if (nTiles && (!dSrc || !dDst)) return fail("xTranspose32x32BatchDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dSrc) | reinterpret_cast<uintptr_t>(dDst)) & 15) return fail("xTranspose32x32BatchDev: 16-byte alignment", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_transpose32(dSrc, dDst, nTiles, (cudaStream_t)stream));
    return 0;
}

extern "C" int xTranspose32x32Batch(const uint8_t* src, uint8_t* dst, size_t nTiles)
{
    if (nTiles && (!src || !dst)) return fail("xTranspose32x32Batch", cudaSuccess);
    if (nTiles == 0) return 0;
    return run_chunked(src, 1024, dst, 1024, nTiles, 16384,
                       [&](void* di, void* dO, size_t n, cudaStream_t st) { return launch_transpose32((const uint8_t*)di, (uint8_t*)dO, n, st); });
}

extern "C" int xIntra32DecideDev(const uint8_t* dCur, const uint8_t* dRefs, uint32_t* dCost, int32_t* dBestMode, size_t n, void* stream)
{
    if (n && (!dCur || !dRefs || !dCost || !dBestMode)) return fail("xIntra32DecideDev", cudaSuccess);
    if (reinterpret_cast<uintptr_t>(dCur) & 3) return fail("xIntra32DecideDev: 4-byte alignment", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_intra32_decide(dCur, dRefs, dCost, dBestMode, n, (cudaStream_t)stream));
    return 0;
}

extern "C" int xIntra32Decide(const uint8_t* cur, const uint8_t* refs, uint32_t* cost, int32_t* bestMode, size_t n)
{
    if (n && (!cur || !refs || !cost || !bestMode)) return fail("xIntra32Decide", cudaSuccess);
    if (n == 0) return 0;
    const HostArr ins[2] = { { const_cast<uint8_t*>(cur), 1024 }, { const_cast<uint8_t*>(refs), 129 } };
    const HostArr outs[2] = { { cost, 35 * 4 }, { bestMode, 4 } };
    return run_chunked(ins, 2, outs, 2, n, (size_t)1 << 14,
                       [&](void* const* dI, void* const* dO, size_t, size_t np, cudaStream_t st) {
                           return launch_intra32_decide((const uint8_t*)dI[0], (const uint8_t*)dI[1], (uint32_t*)dO[0], (int32_t*)dO[1], np, st);
                       });
}

// ---- "next" rows N1 + N3: the closed intra block loop ----------------------------------------------------
extern "C" int xIntra32EncodeBlockDev(const uint8_t* dCur, const uint8_t* dRefs, size_t n, int qp, int16_t* dLevel, uint8_t* dRecon,
                                      int32_t* dBestMode, uint32_t* dCost, void* stream)
{
    if (qp < 0 || qp > 51 || (n && (!dCur || !dRefs || !dLevel || !dRecon || !dBestMode))) return fail("xIntra32EncodeBlockDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dCur) & 7) || (reinterpret_cast<uintptr_t>(dLevel) & 15) || (reinterpret_cast<uintptr_t>(dRecon) & 7))
        return fail("xIntra32EncodeBlockDev: alignment (cur 8, level 16, recon 8 bytes)", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_intra32_encode(dCur, dRefs, n, qp, dLevel, dRecon, dBestMode, dCost, (cudaStream_t)stream));
    return 0;
}

extern "C" int xIntra32EncodeBlock(const uint8_t* cur, const uint8_t* refs, size_t n, int qp, int16_t* level, uint8_t* recon,
                                   int32_t* bestMode, uint32_t* cost)
{
    if (qp < 0 || qp > 51 || (n && (!cur || !refs || !level || !recon || !bestMode))) return fail("xIntra32EncodeBlock", cudaSuccess);
    if (n == 0) return 0;
    const HostArr ins[2] = { { const_cast<uint8_t*>(cur), 1024 }, { const_cast<uint8_t*>(refs), 129 } };
    const HostArr outs[4] = { { level, 2048 }, { recon, 1024 }, { bestMode, 4 }, { cost, 35 * 4 } };
    return run_chunked(ins, 2, outs, 4, n, (size_t)1 << 13,
                       [&](void* const* dI, void* const* dO, size_t, size_t np, cudaStream_t st) {
                           return launch_intra32_encode((const uint8_t*)dI[0], (const uint8_t*)dI[1], np, qp, (int16_t*)dO[0], (uint8_t*)dO[1],
                                                        (int32_t*)dO[2], (uint32_t*)dO[3], st);
                       });
}

extern "C" int xIntra32Recon(const uint8_t* cur, const uint8_t* refs, const uint8_t* mode, size_t n, int qp, int16_t* level, uint8_t* recon)
{
    if (qp < 0 || qp > 51 || (n && (!cur || !refs || !mode || !level || !recon))) return fail("xIntra32Recon", cudaSuccess);
    for (size_t i = 0; i < n; i++)
        if (mode[i] > 34) return fail("xIntra32Recon: mode > 34", cudaSuccess);
    if (n == 0) return 0;
    const HostArr ins[3] = { { const_cast<uint8_t*>(cur), 1024 }, { const_cast<uint8_t*>(refs), 129 }, { const_cast<uint8_t*>(mode), 1 } };
    const HostArr outs[2] = { { level, 2048 }, { recon, 1024 } };
    return run_chunked(ins, 3, outs, 2, n, (size_t)1 << 13,
                       [&](void* const* dI, void* const* dO, size_t, size_t np, cudaStream_t st) {
                           return launch_intra32_recon((const uint8_t*)dI[0], (const uint8_t*)dI[1], (const uint8_t*)dI[2], np, qp,
                                                       (int16_t*)dO[0], (uint8_t*)dO[1], st);
                       });
}

extern "C" int xIntra32ReconDev(const uint8_t* dCur, const uint8_t* dRefs, const uint8_t* dMode, size_t n, int qp, int16_t* dLevel,
                                uint8_t* dRecon, void* stream)
{
    if (qp < 0 || qp > 51 || (n && (!dCur || !dRefs || !dMode || !dLevel || !dRecon))) return fail("xIntra32ReconDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dCur) & 7) || (reinterpret_cast<uintptr_t>(dLevel) & 15) || (reinterpret_cast<uintptr_t>(dRecon) & 7))
        return fail("xIntra32ReconDev: alignment (cur 8, level 16, recon 8 bytes)", cudaSuccess);
    if (dev_ready()) return -1;
    if (check_modes_dev("xIntra32ReconDev", dMode, n, (cudaStream_t)stream)) return -1;
    CK(launch_intra32_recon(dCur, dRefs, dMode, n, qp, dLevel, dRecon, (cudaStream_t)stream));
    return 0;
}

extern "C" int xQuantDequantDev(const int16_t* dCoef, int16_t* dLevel, int16_t* dDequant, size_t nCoef, int qp, void* stream)
{
    if (qp < 0 || qp > 51 || (nCoef & 7) || (nCoef && (!dCoef || (!dLevel && !dDequant)))) return fail("xQuantDequantDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dCoef) | reinterpret_cast<uintptr_t>(dLevel) | reinterpret_cast<uintptr_t>(dDequant)) & 15)
        return fail("xQuantDequantDev: 16-byte alignment", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_quant_dequant(dCoef, dLevel, dDequant, nCoef, qp, (cudaStream_t)stream));
    return 0;
}

// ---- "next" rows: the encoder's tiled frame format on the device ---------------------------------------
extern "C" int xConvInputFmtDev(void* dTiles, const uint8_t* dY, const uint8_t* dU, const uint8_t* dV, intptr_t strdY,
                                int width, int height, void* stream)
{
    // replaces src/x266.cpp:415-453 on device-resident planes
    if (!dTiles || !dY || !dU || !dV || strdY < width || (reinterpret_cast<uintptr_t>(dTiles) & 15)) return fail("xConvInputFmtDev", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_conv_input_fmt((uint8_t*)dTiles, dY, dU, dV, strdY, width, height, (cudaStream_t)stream));
    return 0;
}

extern "C" int xConvOutput420Dev(const void* dTiles, uint8_t* dY, intptr_t strdY, uint8_t* dU, uint8_t* dV, intptr_t strdC,
                                 int width, int height, void* stream)
{
    // replaces src/x266.cpp:455-492
    if (!dTiles || !dY || !dU || !dV || strdY < width || strdC < width / 2 || (reinterpret_cast<uintptr_t>(dTiles) & 15))
        return fail("xConvOutput420Dev", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_conv_output420((const uint8_t*)dTiles, dY, strdY, dU, dV, strdC, width, height, (cudaStream_t)stream));
    return 0;
}

// the reference's own host signatures (src/x266.cpp:415-421, 455-462): void, caller contract violations abort like its asserts
static int conv_input_host(void* pBlock, const uint8_t* inpY, const uint8_t* inpU, const uint8_t* inpV, intptr_t strdY, int width, int height)
{
    if (!pBlock || !inpY || !inpU || !inpV || width <= 0 || height <= 0 || (width & 15) || (height & 15) || strdY < width)
        return fail("xConvInputFmt", cudaSuccess);
    PipeLease l;
    if (!l.ok()) return -1;
    Pipe& p = *l.p;
    const size_t nTiles = (size_t)(width / 16) * (height / 16), lum = (size_t)width * height, chr = lum / 4;
    const intptr_t strdC = strdY >> 1;                                // x266.cpp:426
    if (ensure(&p.dIn[0], &p.capIn[0], lum + 2 * chr)) return -1;
    if (ensure(&p.dOut[0], &p.capOut[0], nTiles * 512)) return -1;
    uint8_t* dY = (uint8_t*)p.dIn[0];
    uint8_t* dU = dY + lum;
    uint8_t* dV = dU + chr;
    cudaStream_t st = p.st[0];
    CK(cudaMemcpy2DAsync(dY, width, inpY, strdY, width, height, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(dU, width / 2, inpU, strdC, width / 2, height / 2, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(dV, width / 2, inpV, strdC, width / 2, height / 2, cudaMemcpyHostToDevice, st));
    CK(launch_conv_input_fmt((uint8_t*)p.dOut[0], dY, dU, dV, width, width, height, st));
    // m_Y | m_C of every tile (384 bytes); m_I stays what the caller had there, as in the reference
    CK(cudaMemcpy2DAsync(pBlock, 512, p.dOut[0], 512, 384, nTiles, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" void xConvInputFmt(ref_block_t* pBlock, const uint8_t* inpY, const uint8_t* inpU, const uint8_t* inpV, const intptr_t strdY,
                              const int width, const int height)
{
    if (conv_input_host(pBlock, inpY, inpU, inpV, strdY, width, height)) die("xConvInputFmt");
}

static int conv_output_host(const void* pBlock, uint8_t* outY, intptr_t strdY, uint8_t* outU, uint8_t* outV, intptr_t strdC, int width, int height)
{
    if (!pBlock || !outY || !outU || !outV || width <= 0 || height <= 0 || (width & 15) || (height & 15) || strdY < width || strdC < width / 2)
        return fail("xConvOutput420", cudaSuccess);
    PipeLease l;
    if (!l.ok()) return -1;
    Pipe& p = *l.p;
    const size_t nTiles = (size_t)(width / 16) * (height / 16), lum = (size_t)width * height, chr = lum / 4;
    if (ensure(&p.dIn[0], &p.capIn[0], nTiles * 512)) return -1;
    if (ensure(&p.dOut[0], &p.capOut[0], lum + 2 * chr)) return -1;
    uint8_t* dY = (uint8_t*)p.dOut[0];
    uint8_t* dU = dY + lum;
    uint8_t* dV = dU + chr;
    cudaStream_t st = p.st[0];
    CK(cudaMemcpyAsync(p.dIn[0], pBlock, nTiles * 512, cudaMemcpyHostToDevice, st));
    CK(launch_conv_output420((const uint8_t*)p.dIn[0], dY, width, dU, dV, width / 2, width, height, st));
    CK(cudaMemcpy2DAsync(outY, strdY, dY, width, width, height, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpy2DAsync(outU, strdC, dU, width / 2, width / 2, height / 2, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpy2DAsync(outV, strdC, dV, width / 2, width / 2, height / 2, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" void xConvOutput420(const ref_block_t* pBlock, uint8_t* outY, const intptr_t strdY, uint8_t* outU, uint8_t* outV, intptr_t strdC,
                               const int width, const int height)
{
    if (conv_output_host(pBlock, outY, strdY, outU, outV, strdC, width, height)) die("xConvOutput420");
}

// Full search on the encoder's own frame stores: current and reference are ref_block_t frames (m_frames[0] and m_frames[1..2],
// src/x266.cpp:99), the reference is edge-replicated by `range` pixels here instead of by the caller.
template <typename CT>
static int search_tiled_dev(const char* api, search_launch_t<CT> launch, const void* dCurTiles, const void* dRefTiles, int w, int h, int range,
                            size_t blk0, size_t blk1, CT* dCost, int32_t* dBest, cudaStream_t st)
{
    if (!dCurTiles || !dRefTiles || w <= 0 || h <= 0 || (w & 15) || (h & 15) || range < 0 || blk1 < blk0 || blk1 > (size_t)(w / 8) * (h / 8))
        return fail(api, cudaSuccess);
    if (blk1 == blk0) return 0;
    if (dev_ready()) return -1;
    const size_t curBytes = align256((size_t)w * h), strd = (size_t)w + 2 * range, refBytes = strd * (h + 2 * range);
    void* planes = nullptr;
    cudaError_t e = scratch_alloc(&planes, curBytes + refBytes, st);
    if (e != cudaSuccess) return fail(api, e);
    uint8_t* dCur = (uint8_t*)planes;
    uint8_t* dRef = dCur + curBytes;
    e = launch_tiles_to_luma((const uint8_t*)dCurTiles, w, h, 0, dCur, st);
    if (e == cudaSuccess) e = launch_tiles_to_luma((const uint8_t*)dRefTiles, w, h, range, dRef, st);
    if (e == cudaSuccess) e = launch(dCur, dRef, (intptr_t)strd, w, h, range, blk0, blk1, dCost, dBest, st);
    const cudaError_t ef = scratch_free(planes, st);
    if (e != cudaSuccess || ef != cudaSuccess) return fail(api, e != cudaSuccess ? e : ef);
    return 0;
}

extern "C" int xSatd8x8SearchTiledDev(const void* dCurTiles, const void* dRefTiles, int width, int height, int range, size_t blk0, size_t blk1,
                                      uint32_t* dCost, int32_t* dBest, void* stream)
{
    return search_tiled_dev<uint32_t>("xSatd8x8SearchTiledDev", launch_satd8x8_search, dCurTiles, dRefTiles, width, height, range, blk0, blk1, dCost, dBest,
                            (cudaStream_t)stream);
}

extern "C" int xSad8x8SearchTiledDev(const void* dCurTiles, const void* dRefTiles, int width, int height, int range, size_t blk0, size_t blk1,
                                     uint32_t* dCost, int32_t* dBest, void* stream)
{
    return search_tiled_dev<uint32_t>("xSad8x8SearchTiledDev", launch_sad8x8_search, dCurTiles, dRefTiles, width, height, range, blk0, blk1, dCost, dBest,
                            (cudaStream_t)stream);
}

extern "C" int xSatd8x8SearchTiledU16Dev(const void* dCurTiles, const void* dRefTiles, int width, int height, int range, size_t blk0, size_t blk1,
                                         uint16_t* dCost, int32_t* dBest, void* stream)
{
    return search_tiled_dev<uint16_t>("xSatd8x8SearchTiledU16Dev", launch_satd8x8_search, dCurTiles, dRefTiles, width, height, range, blk0, blk1, dCost,
                                      dBest, (cudaStream_t)stream);
}

extern "C" int xSad8x8SearchTiledU16Dev(const void* dCurTiles, const void* dRefTiles, int width, int height, int range, size_t blk0, size_t blk1,
                                        uint16_t* dCost, int32_t* dBest, void* stream)
{
    return search_tiled_dev<uint16_t>("xSad8x8SearchTiledU16Dev", launch_sad8x8_search, dCurTiles, dRefTiles, width, height, range, blk0, blk1, dCost,
                                      dBest, (cudaStream_t)stream);
}

// host form: the two frames go up once, the block range runs through the chunked pipeline like xSatd8x8Search
extern "C" int xSatd8x8SearchTiled(const void* curTiles, const void* refTiles, int width, int height, int range, size_t blk0, size_t blk1,
                                   uint32_t* cost, int32_t* best)
{
    const char* api = "xSatd8x8SearchTiled";
    if (!curTiles || !refTiles || width <= 0 || height <= 0 || (width & 15) || (height & 15) || range < 0 || blk1 < blk0 ||
        blk1 > (size_t)(width / 8) * (height / 8))
        return fail(api, cudaSuccess);
    if (blk1 == blk0) return 0;
    PipeLease l;
    if (!l.ok()) return -1;
    Pipe& p = *l.p;
    const size_t tileBytes = (size_t)(width / 16) * (height / 16) * 512;
    const size_t curBytes = align256((size_t)width * height), strd = (size_t)width + 2 * range, refBytes = align256(strd * (height + 2 * range));
    if (ensure(&p.dAux, &p.capAux, 2 * tileBytes + curBytes + refBytes)) return -1;
    uint8_t* dCurT = (uint8_t*)p.dAux;
    uint8_t* dRefT = dCurT + tileBytes;
    uint8_t* dCur = dRefT + tileBytes;
    uint8_t* dRef = dCur + curBytes;
    CK(cudaMemcpyAsync(dCurT, curTiles, tileBytes, cudaMemcpyHostToDevice, p.st[0]));
    CK(cudaMemcpyAsync(dRefT, refTiles, tileBytes, cudaMemcpyHostToDevice, p.st[0]));
    CK(launch_tiles_to_luma(dCurT, width, height, 0, dCur, p.st[0]));
    CK(launch_tiles_to_luma(dRefT, width, height, range, dRef, p.st[0]));
    CK(cudaStreamSynchronize(p.st[0]));
    const size_t side = (size_t)(2 * range + 1), costUnit = side * side * 4;
    size_t per = cost ? ((size_t)64 << 20) / costUnit : (size_t)1 << 20;
    if (per < 1) per = 1;
    const HostArr outs[2] = { { cost, costUnit }, { best, 12 } };
    if (run_chunked_on(p, nullptr, 0, outs, 2, blk1 - blk0, per,
                       [&](void* const*, void* const* dO, size_t u0, size_t nu, cudaStream_t st) {
                           return launch_satd8x8_search(dCur, dRef, (intptr_t)strd, width, height, range, blk0 + u0, blk0 + u0 + nu,
                                                        (uint32_t*)dO[0], (int32_t*)dO[1], st);
                       })) {
        const std::string inner(t_err);
        snprintf(t_err, sizeof(t_err), "%s: %s", api, inner.c_str());
        return -1;
    }
    return 0;
}

extern "C" int xFrameResiDct32Dev(const void* dCurTiles, const void* dPredTiles, int width, int height, int16_t* dCoef,
                                  int s1, int s2, void* stream)
{
    if (!dCurTiles || !dPredTiles || !dCoef || !shifts_ok(s1, s2)) return fail("xFrameResiDct32Dev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dCurTiles) | reinterpret_cast<uintptr_t>(dPredTiles) | reinterpret_cast<uintptr_t>(dCoef)) & 15)
        return fail("xFrameResiDct32Dev: 16-byte alignment", cudaSuccess);
    if (dev_ready()) return -1;
    CK(launch_frame_resi_dct32((const uint8_t*)dCurTiles, (const uint8_t*)dPredTiles, width, height, dCoef, s1, s2, (cudaStream_t)stream));
    return 0;
}

extern "C" int xFrameResiDct32(const void* curTiles, const void* predTiles, int width, int height, int16_t* coef, int s1, int s2)
{
    if (!curTiles || !predTiles || !coef || !shifts_ok(s1, s2) || width <= 0 || height <= 0 || (width & 31) || (height & 31))
        return fail("xFrameResiDct32", cudaSuccess);
    PipeLease l;
    if (!l.ok()) return -1;
    Pipe& p = *l.p;
    const size_t tileBytes = (size_t)(width / 16) * (height / 16) * 512;
    const size_t coefBytes = (size_t)width * height * 2;
    if (ensure(&p.dAux, &p.capAux, 2 * tileBytes)) return -1;
    if (ensure(&p.dOut[0], &p.capOut[0], coefBytes)) return -1;
    uint8_t* dCur = (uint8_t*)p.dAux;
    uint8_t* dPred = dCur + tileBytes;
    CK(cudaMemcpyAsync(dCur, curTiles, tileBytes, cudaMemcpyHostToDevice, p.st[0]));
    CK(cudaMemcpyAsync(dPred, predTiles, tileBytes, cudaMemcpyHostToDevice, p.st[0]));
    CK(launch_frame_resi_dct32(dCur, dPred, width, height, (int16_t*)p.dOut[0], s1, s2, p.st[0]));
    CK(cudaMemcpyAsync(coef, p.dOut[0], coefBytes, cudaMemcpyDeviceToHost, p.st[0]));
    CK(cudaStreamSynchronize(p.st[0]));
    return 0;
}

// ================================================================================================
// Tier 2 (reference signatures; host pointers; synchronous)
// ================================================================================================
extern "C" void partialButterfly32(const int16_t* src, int16_t* dst, int shift, int line)
{
    // replaces src_tb/dct32.c:66-170
    if (line <= 0) return;
    const size_t bytes = (size_t)line * 64;
    auto body = [&]() -> int {
        PipeLease l;
        if (!l.ok()) return -1;
        Pipe& p = *l.p;
        if (ensure(&p.dIn[0], &p.capIn[0], bytes)) return -1;
        if (ensure(&p.dOut[0], &p.capOut[0], bytes)) return -1;
        CK(cudaMemcpyAsync(p.dIn[0], src, bytes, cudaMemcpyHostToDevice, p.st[0]));
        CK(launch_partial32((const int16_t*)p.dIn[0], (int16_t*)p.dOut[0], shift, line, p.st[0]));
        CK(cudaMemcpyAsync(dst, p.dOut[0], bytes, cudaMemcpyDeviceToHost, p.st[0]));
        CK(cudaStreamSynchronize(p.st[0]));
        return 0;
    };
    if (shift < 1 || shift > 16) { fail("partialButterfly32: shift", cudaSuccess); die("partialButterfly32"); }
    if (body()) die("partialButterfly32");
}

extern "C" int satd8x8(const int16_t diff[64])
{
    // replaces src_tb/satd.c:31-118
    int32_t out = 0;
    if (xSatd8x8Batch(diff, &out, 1)) die("satd8x8");
    return out;
}

// ================================================================================================
// Tier 1 (BDPI stream API of the golden model; module-global state like the reference)
// ================================================================================================
static int16_t s_dctMat[32 * 32];      // dct32.c:173
static int16_t s_dctOut[32 * 32];      // dct32.c:174
static int s_lastDiff = 0, s_lastDct = 0;

extern "C" void dct32_genNew(void)
{
    // stimulus exactly as src_tb/dct32.c:187-195; the two passes (:197-198) run on the GPU
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            const int a = rand() & 0xFF;
            const int b = rand() & 0xFF;
            s_dctMat[i * 32 + j] = (int16_t)(a - b);
        }
    if (xDct32Batch(s_dctMat, s_dctOut, 1, 4, 11)) die("dct32_genNew");
    s_lastDiff = 0;
    s_lastDct = 0;
}

extern "C" void dct32_getDiff(unsigned int res[])
{
    // src_tb/dct32.c:205-220: two rows, two samples per little-endian word
    int x = 0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 32; j += 2) {
            const unsigned lo = (unsigned short)s_dctMat[(s_lastDiff + i) * 32 + j];
            const unsigned hi = (unsigned short)s_dctMat[(s_lastDiff + i) * 32 + j + 1];
            res[x++] = (hi << 16) + lo;
        }
    s_lastDiff += 2;
}

extern "C" unsigned long long dct32_getDct(void)
{
    // src_tb/dct32.c:223-246: 4 vertically adjacent coefficients of one column per call
    const int col = s_lastDct >> 5, row = s_lastDct & 31;
    unsigned long long ret = 0;
    for (int i = 0; i < 4; i++)
        ret |= (unsigned long long)(unsigned short)s_dctOut[(row + i) * 32 + col] << (16 * i);
    s_lastDct += 4;
    return ret;
}

static int16_t s_satdMat[8 * 8];       // satd.c:120
static int s_lastRow = 0;
static int s_lastSatd = 0;

extern "C" void satd8x8_genNew(void)
{
    // stimulus as src_tb/satd.c:128-136; cost (:138) from the GPU kernel
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            const int a = rand() & 0xFF;
            const int b = rand() & 0xFF;
            s_satdMat[i * 8 + j] = (int16_t)(a - b);
        }
    s_lastSatd = satd8x8(s_satdMat);
    s_lastRow = 0;
}

extern "C" void satd8x8_getDiff(unsigned int res[])
{
    memcpy(res, &s_satdMat[s_lastRow * 8], 8 * sizeof(int16_t));   // satd.c:143-147
    s_lastRow++;
}

extern "C" unsigned int satd8x8_getSatd(void) { return (unsigned int)s_lastSatd; }   // satd.c:149-152
