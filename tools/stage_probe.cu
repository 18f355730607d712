// stage_probe.cu -- does a CACHE-RESIDENT pinned staging ring save host DRAM traffic on this box?
// One direction at a time, no kernel: pageable source -> pinned slot (pool copy, NT or cached stores) -> cudaMemcpyAsync H2D, and
// D2H -> pinned slot -> pageable destination, over rings of 4 slots of 0.5..32 MiB.  If the DMA engine could read cached-store
// data out of the last-level cache, small rings with cached stores would beat large rings with non-temporal stores.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include "../x266_b200/csrc/hostcopy.h"
using namespace x266;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main()
{
    const size_t total = (size_t)1 << 30;
    char* src = (char*)aligned_alloc(4096, total);
    char* dst = (char*)aligned_alloc(4096, total);
    memset(src, 1, total); memset(dst, 2, total);
    char* dev; CK(cudaMalloc(&dev, total));
    cudaStream_t st[4]; cudaEvent_t ev[4];
    for (int i = 0; i < 4; i++) { CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); }
    set_host_copy_threads(8);
    for (size_t slot : { (size_t)512 << 10, (size_t)1 << 20, (size_t)2 << 20, (size_t)4 << 20, (size_t)8 << 20, (size_t)32 << 20 }) {
        char* pin[4];
        for (int i = 0; i < 4; i++) CK(cudaHostAlloc(&pin[i], slot, cudaHostAllocDefault));
        for (int nt = 0; nt <= 3; nt += 3) {
            set_host_copy_nt(nt);
            for (int dir = 0; dir < 2; dir++) {
                const size_t n = total / slot;
                auto t0 = std::chrono::steady_clock::now();
                for (size_t i = 0; i < n; i++) {
                    const int s = (int)(i & 3);
                    if (i >= 4) CK(cudaEventSynchronize(ev[s]));
                    if (dir == 0) {
                        CopyJob j{ pin[s], src + i * slot, slot, false };
                        host_copy_parallel(&j, 1);
                        CK(cudaMemcpyAsync(dev + i * slot, pin[s], slot, cudaMemcpyHostToDevice, st[s]));
                        CK(cudaEventRecord(ev[s], st[s]));
                    } else {
                        // D2H of chunk i into slot s, copy-out of chunk i-2 (LAG 2)
                        CK(cudaMemcpyAsync(pin[s], dev + i * slot, slot, cudaMemcpyDeviceToHost, st[s]));
                        CK(cudaEventRecord(ev[s], st[s]));
                        if (i >= 2) {
                            const int sj = (int)((i - 2) & 3);
                            CK(cudaEventSynchronize(ev[sj]));
                            CopyJob j{ dst + (i - 2) * slot, pin[sj], slot, true };
                            host_copy_parallel(&j, 1);
                        }
                    }
                }
                CK(cudaDeviceSynchronize());
                const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                printf("slot %5zu KiB  %s  %s: %6.1f GB/s\n", slot >> 10, nt ? "NT stores    " : "cached stores", dir ? "D2H + copy-out" : "copy-in + H2D ", total / dt / 1e9);
                fflush(stdout);
            }
        }
        for (int i = 0; i < 4; i++) cudaFreeHost(pin[i]);
    }
    return 0;
}
