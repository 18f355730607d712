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

#include <atomic>
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
constexpr int SLOTS = 3;

struct Ctx {
    bool ready = false;
    int dev = -1;
    int sms = 0;
    cudaStream_t st[SLOTS] = {};
    void* dIn[SLOTS] = {};
    void* dOut[SLOTS] = {};
    size_t capIn[SLOTS] = {};
    size_t capOut[SLOTS] = {};
    void* dAux = nullptr;          // persistent inputs shared by all chunks of one call (search planes)
    size_t capAux = 0;
    std::mutex mu;                 // serialises host-pointer calls on this device
};

static Ctx g_ctx[MAX_DEV];
static std::mutex g_initMu;
static int g_smCache[MAX_DEV] = {};

int sm_count()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return 148;
    if (!g_smCache[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_smCache[dev] = n;
    }
    return g_smCache[dev];
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

static cudaMemPool_t g_pool[MAX_DEV] = {};
static std::mutex g_poolMu;

cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEV) return cudaErrorInvalidDevice;
    if (!g_pool[dev]) {
        std::lock_guard<std::mutex> lk(g_poolMu);
        if (!g_pool[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t pool;
            if ((e = cudaMemPoolCreate(&pool, &props)) != cudaSuccess) return e;
            unsigned long long keep = ~0ull;
            if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
            g_pool[dev] = pool;
        }
    }
    return cudaMallocFromPoolAsync(p, bytes, g_pool[dev], st);
}

cudaError_t scratch_free(void* p, cudaStream_t st) { return cudaFreeAsync(p, st); }

static int ctx_get(Ctx** out)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) return fail("device index", cudaSuccess);
    Ctx& c = g_ctx[dev];
    if (!c.ready) {
        std::lock_guard<std::mutex> lk(g_initMu);
        if (!c.ready) {
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, dev));
            if (prop.major != 10) {
                snprintf(t_err, sizeof(t_err), "device %d is sm_%d%d; libx266_b200 is built for sm_100a only", dev, prop.major, prop.minor);
                return -1;
            }
            c.dev = dev;
            c.sms = prop.multiProcessorCount;
            for (int i = 0; i < SLOTS; i++) CK(cudaStreamCreateWithFlags(&c.st[i], cudaStreamNonBlocking));
            c.ready = true;
        }
    }
    *out = &c;
    return 0;
}

static int ensure(void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Chunked pipeline: units are independent, so chunk i runs H2D -> kernel -> D2H on stream i%3; the
// three streams overlap the two copy directions with compute.  Stream order makes slot reuse safe.
// ------------------------------------------------------------------------------------------------
template <typename Launch>
static int run_chunked(Ctx& c, const void* src, size_t inUnit, void* dst, size_t outUnit, size_t nUnits,
                       size_t unitsPerChunk, Launch launch)
{
    std::lock_guard<std::mutex> lk(c.mu);
    const size_t chunk = nUnits < unitsPerChunk ? nUnits : unitsPerChunk;
    size_t i = 0;
    for (size_t u0 = 0; u0 < nUnits; u0 += chunk, i++) {
        const int s = (int)(i % SLOTS);
        const size_t nu = (nUnits - u0) < chunk ? (nUnits - u0) : chunk;
        if (ensure(&c.dIn[s], &c.capIn[s], chunk * inUnit)) return -1;
        if (ensure(&c.dOut[s], &c.capOut[s], chunk * outUnit)) return -1;
        CK(cudaMemcpyAsync(c.dIn[s], (const char*)src + u0 * inUnit, nu * inUnit, cudaMemcpyHostToDevice, c.st[s]));
        CK(launch(c.dIn[s], c.dOut[s], nu, c.st[s]));
        CK(cudaMemcpyAsync((char*)dst + u0 * outUnit, c.dOut[s], nu * outUnit, cudaMemcpyDeviceToHost, c.st[s]));
    }
    for (int s = 0; s < SLOTS; s++) CK(cudaStreamSynchronize(c.st[s]));
    return 0;
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
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return;
    Ctx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lk(g_initMu);
    {
        std::lock_guard<std::mutex> lp(g_poolMu);
        if (g_pool[dev]) {                       // scratch of the search kernels: caller has synchronised its streams
            cudaDeviceSynchronize();
            cudaMemPoolDestroy(g_pool[dev]);
            g_pool[dev] = nullptr;
        }
    }
    if (!c.ready) return;
    for (int i = 0; i < SLOTS; i++) {
        cudaStreamSynchronize(c.st[i]);
        cudaStreamDestroy(c.st[i]);
        if (c.dIn[i]) cudaFree(c.dIn[i]);
        if (c.dOut[i]) cudaFree(c.dOut[i]);
        c.dIn[i] = c.dOut[i] = nullptr; c.capIn[i] = c.capOut[i] = 0;
    }
    if (c.dAux) cudaFree(c.dAux);
    c.dAux = nullptr; c.capAux = 0;
    c.ready = false;
}

extern "C" const char* xGpuLastError(void) { return t_err; }

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
    return fail("xGpuTune: unknown key", cudaSuccess);
}

extern "C" int xDct32BatchDev(const int16_t* dSrc, int16_t* dDst, size_t nBlocks, int s1, int s2, void* stream)
{
    if (!shifts_ok(s1, s2) || (nBlocks && (!dSrc || !dDst))) return fail("xDct32BatchDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dSrc) | reinterpret_cast<uintptr_t>(dDst)) & 15) return fail("xDct32BatchDev: 16-byte alignment", cudaSuccess);
    CK(dct32_dispatch(dSrc, dDst, nBlocks, s1, s2, (cudaStream_t)stream));
    return 0;
}

extern "C" int xDct32Batch(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2)
{
    if (!shifts_ok(s1, s2) || (nBlocks && (!src || !dst))) return fail("xDct32Batch", cudaSuccess);
    if (nBlocks == 0) return 0;
    Ctx* c;
    if (ctx_get(&c)) return -1;
    // 32 MiB chunks amortise the per-chunk hand-over best (47.0 GB/s each way of the 48.2 the link gives with both directions busy);
    // below ~24 chunks the fill and drain of the three-stage pipeline cost more than that, so smaller batches use 16 Ki blocks
    // (profiles/r01_e2e_chunk_sweep.log).
    size_t chunk = g_dctChunk.load();
    if (chunk == 0) chunk = nBlocks >= ((size_t)3 << 18) ? 32768 : 16384;
    return run_chunked(*c, src, 2048, dst, 2048, nBlocks, chunk,
                       [&](void* di, void* dO, size_t n, cudaStream_t st) {
                           return dct32_dispatch((const int16_t*)di, (int16_t*)dO, n, s1, s2, st);
                       });
}

extern "C" int xDct32BatchMultiGpu(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, int nGpus)
{
    // blocks are independent -> GPU g of G owns the contiguous range [g*N/G, (g+1)*N/G); one host
    // thread per device drives that device's chunked pipeline; no collective, no peer traffic.
    int have = 0;
    CK(cudaGetDeviceCount(&have));
    if (nGpus <= 0) nGpus = have;
    if (nGpus > have || !shifts_ok(s1, s2) || (nBlocks && (!src || !dst))) return fail("xDct32BatchMultiGpu", cudaSuccess);
    if (nBlocks == 0) return 0;
    int prev = 0;
    CK(cudaGetDevice(&prev));
    std::vector<int> rc(nGpus, 0);
    std::vector<std::string> err(nGpus);
    std::vector<std::thread> th;
    for (int g = 0; g < nGpus; g++)
        th.emplace_back([&, g]() {
            const size_t lo = nBlocks * (size_t)g / nGpus, hi = nBlocks * (size_t)(g + 1) / nGpus;
            if (cudaSetDevice(g) != cudaSuccess) { rc[g] = -1; err[g] = "cudaSetDevice failed"; return; }
            rc[g] = xDct32Batch(src + lo * 1024, dst + lo * 1024, hi - lo, s1, s2);
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
    CK(launch_idct32_imma(dSrc, dDst, nBlocks, s1, s2, (cudaStream_t)stream));
    return 0;
}

extern "C" int xIdct32Batch(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2)
{
    if (!shifts_ok(s1, s2) || (nBlocks && (!src || !dst))) return fail("xIdct32Batch", cudaSuccess);
    if (nBlocks == 0) return 0;
    Ctx* c;
    if (ctx_get(&c)) return -1;
    return run_chunked(*c, src, 2048, dst, 2048, nBlocks, 8192,
                       [&](void* di, void* dO, size_t n, cudaStream_t st) {
                           return launch_idct32_imma((const int16_t*)di, (int16_t*)dO, n, s1, s2, st);
                       });
}

extern "C" int xDctNBatchDev(int log2N, const int16_t* dSrc, int16_t* dDst, size_t nBlocks, int s1, int s2, void* stream)
{
    if (log2N == 5) return xDct32BatchDev(dSrc, dDst, nBlocks, s1, s2, stream);
    if (log2N < 2 || log2N > 5 || !shifts_ok(s1, s2) || (nBlocks && (!dSrc || !dDst))) return fail("xDctNBatchDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dSrc) | reinterpret_cast<uintptr_t>(dDst)) & 15) return fail("xDctNBatchDev: 16-byte alignment", cudaSuccess);
    CK(launch_dctN(log2N, dSrc, dDst, nBlocks, s1, s2, (cudaStream_t)stream));
    return 0;
}

extern "C" int xDctNBatch(int log2N, const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2)
{
    if (log2N == 5) return xDct32Batch(src, dst, nBlocks, s1, s2);
    if (log2N < 2 || log2N > 5 || !shifts_ok(s1, s2) || (nBlocks && (!src || !dst))) return fail("xDctNBatch", cudaSuccess);
    if (nBlocks == 0) return 0;
    Ctx* c;
    if (ctx_get(&c)) return -1;
    const size_t unit = (size_t)2 << (2 * log2N);
    return run_chunked(*c, src, unit, dst, unit, nBlocks, (size_t)(16u << 20) / unit,
                       [&](void* di, void* dO, size_t n, cudaStream_t st) {
                           return launch_dctN(log2N, (const int16_t*)di, (int16_t*)dO, n, s1, s2, st);
                       });
}

extern "C" int xPartialButterfly32Dev(const int16_t* dSrc, int16_t* dDst, int shift, int line, void* stream)
{
    if (shift < 1 || shift > 16 || line < 0 || (line && (!dSrc || !dDst))) return fail("xPartialButterfly32Dev", cudaSuccess);
    CK(launch_partial32(dSrc, dDst, shift, line, (cudaStream_t)stream));
    return 0;
}

extern "C" int xSatd8x8BatchDev(const int16_t* dDiff, int32_t* dSatd, size_t n, void* stream)
{
    if (n && (!dDiff || !dSatd)) return fail("xSatd8x8BatchDev", cudaSuccess);
    if (reinterpret_cast<uintptr_t>(dDiff) & 15) return fail("xSatd8x8BatchDev: 16-byte alignment", cudaSuccess);
    CK(launch_satd8x8_batch(dDiff, dSatd, n, (cudaStream_t)stream));
    return 0;
}

extern "C" int xSatd8x8Batch(const int16_t* diff, int32_t* satd, size_t n)
{
    if (n && (!diff || !satd)) return fail("xSatd8x8Batch", cudaSuccess);
    if (n == 0) return 0;
    Ctx* c;
    if (ctx_get(&c)) return -1;
    return run_chunked(*c, diff, 128, satd, 4, n, (size_t)1 << 17,
                       [&](void* di, void* dO, size_t m, cudaStream_t st) {
                           return launch_satd8x8_batch((const int16_t*)di, (int32_t*)dO, m, st);
                       });
}

extern "C" int xSatd8x8SearchDev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                                 size_t blk0, size_t blk1, uint32_t* dCost, int32_t* dBest, void* stream)
{
    if (!dCur || !dRefPadded || w <= 0 || h <= 0 || strd < w + 2 * range) return fail("xSatd8x8SearchDev", cudaSuccess);
    CK(launch_satd8x8_search(dCur, dRefPadded, strd, w, h, range, blk0, blk1, dCost, dBest, (cudaStream_t)stream));
    return 0;
}

typedef cudaError_t (*search_launch_t)(const uint8_t*, const uint8_t*, intptr_t, int, int, int, size_t, size_t, uint32_t*, int32_t*, cudaStream_t);
static int search_host(search_launch_t launch, const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                       size_t blk0, size_t blk1, uint32_t* cost, int32_t* best);

extern "C" int xSatd8x8Search(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                              size_t blk0, size_t blk1, uint32_t* cost, int32_t* best)
{
    return search_host(launch_satd8x8_search, cur, refPadded, strd, w, h, range, blk0, blk1, cost, best);
}

extern "C" int xSad8x8Search(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                             size_t blk0, size_t blk1, uint32_t* cost, int32_t* best)
{
    return search_host(launch_sad8x8_search, cur, refPadded, strd, w, h, range, blk0, blk1, cost, best);
}

extern "C" int xSad8x8SearchDev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                                size_t blk0, size_t blk1, uint32_t* dCost, int32_t* dBest, void* stream)
{
    if (!dCur || !dRefPadded || w <= 0 || h <= 0 || strd < w + 2 * range) return fail("xSad8x8SearchDev", cudaSuccess);
    CK(launch_sad8x8_search(dCur, dRefPadded, strd, w, h, range, blk0, blk1, dCost, dBest, (cudaStream_t)stream));
    return 0;
}

extern "C" int sad(unsigned char* input_data1, unsigned char* input_data2, size_t n)
{
    // replaces riscv/programs/benchmarks/sad/sad.c:27-38 (host pointers, synchronous)
    Ctx* c;
    if (ctx_get(&c)) die("sad");
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t bytes = n * n;
    unsigned out = 0;
    auto body = [&]() -> int {
        if (ensure(&c->dIn[0], &c->capIn[0], 2 * bytes + 256)) return -1;
        if (ensure(&c->dOut[0], &c->capOut[0], 256)) return -1;
        uint8_t* dA = (uint8_t*)c->dIn[0];
        uint8_t* dB = dA + ((bytes + 255) & ~(size_t)255);
        CK(cudaMemcpyAsync(dA, input_data1, bytes, cudaMemcpyHostToDevice, c->st[0]));
        CK(cudaMemcpyAsync(dB, input_data2, bytes, cudaMemcpyHostToDevice, c->st[0]));
        CK(launch_sad_region(dA, dB, bytes, (unsigned*)c->dOut[0], c->st[0]));
        CK(cudaMemcpyAsync(&out, c->dOut[0], sizeof(out), cudaMemcpyDeviceToHost, c->st[0]));
        CK(cudaStreamSynchronize(c->st[0]));
        return 0;
    };
    if (body()) die("sad");
    return (int)out;
}

static int search_host(search_launch_t launch, const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                       size_t blk0, size_t blk1, uint32_t* cost, int32_t* best)
{
    if (!cur || !refPadded || w <= 0 || h <= 0 || (w & 7) || (h & 7) || range < 0 || strd < w + 2 * range || blk1 < blk0 ||
        blk1 > (size_t)(w / 8) * (h / 8))
        return fail("xSatd8x8Search", cudaSuccess);
    if (blk1 == blk0) return 0;
    Ctx* c;
    if (ctx_get(&c)) return -1;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t curBytes = (size_t)w * h;
    const size_t refBytes = (size_t)strd * (h + 2 * range);
    const size_t refOff = (curBytes + 255) & ~(size_t)255;
    if (ensure(&c->dAux, &c->capAux, refOff + refBytes)) return -1;
    uint8_t* dCur = (uint8_t*)c->dAux;
    uint8_t* dRef = dCur + refOff;
    CK(cudaMemcpyAsync(dCur, cur, curBytes, cudaMemcpyHostToDevice, c->st[0]));
    CK(cudaMemcpyAsync(dRef, refPadded, refBytes, cudaMemcpyHostToDevice, c->st[0]));
    CK(cudaStreamSynchronize(c->st[0]));
    const size_t side = (size_t)(2 * range + 1);
    const size_t costUnit = cost ? side * side * 4 : 0;
    // blocks per chunk: keep a chunk's cost surface around 64 MiB
    size_t per = cost ? ((size_t)64 << 20) / costUnit : (size_t)1 << 20;
    if (per < 1) per = 1;
    size_t i = 0;
    for (size_t b0 = blk0; b0 < blk1; b0 += per, i++) {
        const int s = (int)(i % SLOTS);
        const size_t nb = (blk1 - b0) < per ? (blk1 - b0) : per;
        uint32_t* dCost = nullptr;
        int32_t* dBest = nullptr;
        if (cost) { if (ensure(&c->dOut[s], &c->capOut[s], per * costUnit)) return -1; dCost = (uint32_t*)c->dOut[s]; }
        if (best) { if (ensure(&c->dIn[s], &c->capIn[s], per * 12)) return -1; dBest = (int32_t*)c->dIn[s]; }
        CK(launch(dCur, dRef, strd, w, h, range, b0, b0 + nb, dCost, dBest, c->st[s]));
        if (cost) CK(cudaMemcpyAsync(cost + (b0 - blk0) * side * side, dCost, nb * costUnit, cudaMemcpyDeviceToHost, c->st[s]));
        if (best) CK(cudaMemcpyAsync(best + (b0 - blk0) * 3, dBest, nb * 12, cudaMemcpyDeviceToHost, c->st[s]));
    }
    for (int s = 0; s < SLOTS; s++) CK(cudaStreamSynchronize(c->st[s]));
    return 0;
}

extern "C" int xIntra32PredDev(const uint8_t* dRefs, const uint8_t* dMode, uint8_t* dPred, size_t n, void* stream)
{
    if (n && (!dRefs || !dMode || !dPred)) return fail("xIntra32PredDev", cudaSuccess);
    if (reinterpret_cast<uintptr_t>(dPred) & 15) return fail("xIntra32PredDev: 16-byte alignment of pred", cudaSuccess);
    CK(launch_intra32(dRefs, dMode, dPred, n, (cudaStream_t)stream));
    return 0;
}

extern "C" int xIntra32Pred(const uint8_t* refs, const uint8_t* mode, uint8_t* pred, size_t n)
{
    if (n && (!refs || !mode || !pred)) return fail("xIntra32Pred", cudaSuccess);
    for (size_t i = 0; i < n; i++)
        if (mode[i] > 34) return fail("xIntra32Pred: mode > 34", cudaSuccess);
    if (n == 0) return 0;
    Ctx* c;
    if (ctx_get(&c)) return -1;
    // inputs are 129 + 1 bytes per prediction: pack refs and mode into one staged unit stream
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t per = (size_t)1 << 15;
    size_t i = 0;
    for (size_t p0 = 0; p0 < n; p0 += per, i++) {
        const int s = (int)(i % SLOTS);
        const size_t np = (n - p0) < per ? (n - p0) : per;
        if (ensure(&c->dIn[s], &c->capIn[s], per * 130 + 256)) return -1;
        if (ensure(&c->dOut[s], &c->capOut[s], per * 1024)) return -1;
        uint8_t* dRefs = (uint8_t*)c->dIn[s];
        uint8_t* dMode = dRefs + ((per * 129 + 255) & ~(size_t)255);
        CK(cudaMemcpyAsync(dRefs, refs + p0 * 129, np * 129, cudaMemcpyHostToDevice, c->st[s]));
        CK(cudaMemcpyAsync(dMode, mode + p0, np, cudaMemcpyHostToDevice, c->st[s]));
        CK(launch_intra32(dRefs, dMode, (uint8_t*)c->dOut[s], np, c->st[s]));
        CK(cudaMemcpyAsync(pred + p0 * 1024, c->dOut[s], np * 1024, cudaMemcpyDeviceToHost, c->st[s]));
    }
    for (int s = 0; s < SLOTS; s++) CK(cudaStreamSynchronize(c->st[s]));
    return 0;
}

extern "C" int xTranspose32x32BatchDev(const uint8_t* dSrc, uint8_t* dDst, size_t nTiles, void* stream)
{
    if (nTiles && (!dSrc || !dDst)) return fail("xTranspose32x32BatchDev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dSrc) | reinterpret_cast<uintptr_t>(dDst)) & 15) return fail("xTranspose32x32BatchDev: 16-byte alignment", cudaSuccess);
    CK(launch_transpose32(dSrc, dDst, nTiles, (cudaStream_t)stream));
    return 0;
}

extern "C" int xTranspose32x32Batch(const uint8_t* src, uint8_t* dst, size_t nTiles)
{
    if (nTiles && (!src || !dst)) return fail("xTranspose32x32Batch", cudaSuccess);
    if (nTiles == 0) return 0;
    Ctx* c;
    if (ctx_get(&c)) return -1;
    return run_chunked(*c, src, 1024, dst, 1024, nTiles, 16384,
                       [&](void* di, void* dO, size_t n, cudaStream_t st) { return launch_transpose32((const uint8_t*)di, (uint8_t*)dO, n, st); });
}

extern "C" int xIntra32DecideDev(const uint8_t* dCur, const uint8_t* dRefs, uint32_t* dCost, int32_t* dBestMode, size_t n, void* stream)
{
    if (n && (!dCur || !dRefs || !dCost || !dBestMode)) return fail("xIntra32DecideDev", cudaSuccess);
    if (reinterpret_cast<uintptr_t>(dCur) & 3) return fail("xIntra32DecideDev: 4-byte alignment", cudaSuccess);
    CK(launch_intra32_decide(dCur, dRefs, dCost, dBestMode, n, (cudaStream_t)stream));
    return 0;
}

extern "C" int xIntra32Decide(const uint8_t* cur, const uint8_t* refs, uint32_t* cost, int32_t* bestMode, size_t n)
{
    if (n && (!cur || !refs || !cost || !bestMode)) return fail("xIntra32Decide", cudaSuccess);
    if (n == 0) return 0;
    Ctx* c;
    if (ctx_get(&c)) return -1;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t per = (size_t)1 << 14;
    size_t i = 0;
    for (size_t p0 = 0; p0 < n; p0 += per, i++) {
        const int s = (int)(i % SLOTS);
        const size_t np = (n - p0) < per ? (n - p0) : per;
        const size_t refOff = per * 1024, costOff = 0, bestOff = (per * 35 * 4 + 255) & ~(size_t)255;
        if (ensure(&c->dIn[s], &c->capIn[s], per * 1024 + per * 129 + 256)) return -1;
        if (ensure(&c->dOut[s], &c->capOut[s], bestOff + per * 4)) return -1;
        uint8_t* dCur = (uint8_t*)c->dIn[s];
        uint8_t* dRefs = dCur + refOff;
        uint32_t* dCost = (uint32_t*)((uint8_t*)c->dOut[s] + costOff);
        int32_t* dBest = (int32_t*)((uint8_t*)c->dOut[s] + bestOff);
        CK(cudaMemcpyAsync(dCur, cur + p0 * 1024, np * 1024, cudaMemcpyHostToDevice, c->st[s]));
        CK(cudaMemcpyAsync(dRefs, refs + p0 * 129, np * 129, cudaMemcpyHostToDevice, c->st[s]));
        CK(launch_intra32_decide(dCur, dRefs, dCost, dBest, np, c->st[s]));
        CK(cudaMemcpyAsync(cost + p0 * 35, dCost, np * 35 * 4, cudaMemcpyDeviceToHost, c->st[s]));
        CK(cudaMemcpyAsync(bestMode + p0, dBest, np * 4, cudaMemcpyDeviceToHost, c->st[s]));
    }
    for (int s = 0; s < SLOTS; s++) CK(cudaStreamSynchronize(c->st[s]));
    return 0;
}

// ---- "next" rows: the encoder's tiled frame format on the device ---------------------------------------
extern "C" int xConvInputFmtDev(void* dTiles, const uint8_t* dY, const uint8_t* dU, const uint8_t* dV, intptr_t strdY,
                                int width, int height, void* stream)
{
    // replaces src/x266.cpp:415-453 on device-resident planes
    if (!dTiles || !dY || !dU || !dV || strdY < width || (reinterpret_cast<uintptr_t>(dTiles) & 15)) return fail("xConvInputFmtDev", cudaSuccess);
    CK(launch_conv_input_fmt((uint8_t*)dTiles, dY, dU, dV, strdY, width, height, (cudaStream_t)stream));
    return 0;
}

extern "C" int xConvOutput420Dev(const void* dTiles, uint8_t* dY, intptr_t strdY, uint8_t* dU, uint8_t* dV, intptr_t strdC,
                                 int width, int height, void* stream)
{
    // replaces src/x266.cpp:455-492
    if (!dTiles || !dY || !dU || !dV || strdY < width || strdC < width / 2 || (reinterpret_cast<uintptr_t>(dTiles) & 15))
        return fail("xConvOutput420Dev", cudaSuccess);
    CK(launch_conv_output420((const uint8_t*)dTiles, dY, strdY, dU, dV, strdC, width, height, (cudaStream_t)stream));
    return 0;
}

extern "C" int xFrameResiDct32Dev(const void* dCurTiles, const void* dPredTiles, int width, int height, int16_t* dCoef,
                                  int s1, int s2, void* stream)
{
    if (!dCurTiles || !dPredTiles || !dCoef || !shifts_ok(s1, s2)) return fail("xFrameResiDct32Dev", cudaSuccess);
    if ((reinterpret_cast<uintptr_t>(dCurTiles) | reinterpret_cast<uintptr_t>(dPredTiles) | reinterpret_cast<uintptr_t>(dCoef)) & 15)
        return fail("xFrameResiDct32Dev: 16-byte alignment", cudaSuccess);
    CK(launch_frame_resi_dct32((const uint8_t*)dCurTiles, (const uint8_t*)dPredTiles, width, height, dCoef, s1, s2, (cudaStream_t)stream));
    return 0;
}

extern "C" int xFrameResiDct32(const void* curTiles, const void* predTiles, int width, int height, int16_t* coef, int s1, int s2)
{
    if (!curTiles || !predTiles || !coef || !shifts_ok(s1, s2) || width <= 0 || height <= 0 || (width & 31) || (height & 31))
        return fail("xFrameResiDct32", cudaSuccess);
    Ctx* c;
    if (ctx_get(&c)) return -1;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t tileBytes = (size_t)(width / 16) * (height / 16) * 512;
    const size_t coefBytes = (size_t)width * height * 2;
    if (ensure(&c->dAux, &c->capAux, 2 * tileBytes)) return -1;
    if (ensure(&c->dOut[0], &c->capOut[0], coefBytes)) return -1;
    uint8_t* dCur = (uint8_t*)c->dAux;
    uint8_t* dPred = dCur + tileBytes;
    CK(cudaMemcpyAsync(dCur, curTiles, tileBytes, cudaMemcpyHostToDevice, c->st[0]));
    CK(cudaMemcpyAsync(dPred, predTiles, tileBytes, cudaMemcpyHostToDevice, c->st[0]));
    CK(launch_frame_resi_dct32(dCur, dPred, width, height, (int16_t*)c->dOut[0], s1, s2, c->st[0]));
    CK(cudaMemcpyAsync(coef, c->dOut[0], coefBytes, cudaMemcpyDeviceToHost, c->st[0]));
    CK(cudaStreamSynchronize(c->st[0]));
    return 0;
}

// ================================================================================================
// Tier 2 (reference signatures; host pointers; synchronous)
// ================================================================================================
extern "C" void partialButterfly32(const int16_t* src, int16_t* dst, int shift, int line)
{
    // replaces src_tb/dct32.c:66-170
    if (line <= 0) return;
    Ctx* c;
    if (ctx_get(&c)) die("partialButterfly32");
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t bytes = (size_t)line * 64;
    auto body = [&]() -> int {
        if (ensure(&c->dIn[0], &c->capIn[0], bytes)) return -1;
        if (ensure(&c->dOut[0], &c->capOut[0], bytes)) return -1;
        CK(cudaMemcpyAsync(c->dIn[0], src, bytes, cudaMemcpyHostToDevice, c->st[0]));
        CK(launch_partial32((const int16_t*)c->dIn[0], (int16_t*)c->dOut[0], shift, line, c->st[0]));
        CK(cudaMemcpyAsync(dst, c->dOut[0], bytes, cudaMemcpyDeviceToHost, c->st[0]));
        CK(cudaStreamSynchronize(c->st[0]));
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
