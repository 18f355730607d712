// hostcopy.h -- host-side staging copies of the pageable-buffer path (ffi.cu: run_chunked).
// The reference's caller hands over plain _aligned_malloc memory (src/x266.cpp:505,647-649); DMA engines cannot read it,
// so every chunk travels caller memory <-> library-owned pinned ring by CPU copies.  Those copies are the host-side
// bound of that path, so they are spread over a small pool of threads and written with non-temporal stores.
#pragma once
#include <stddef.h>

namespace x266 {

struct CopyJob { void* dst; const void* src; size_t bytes; bool toCaller; };   // toCaller: pinned slot -> caller memory (else caller -> pinned slot)

// Runs all jobs (split into slices, shared over the pool and the calling thread); returns when every byte is copied.
// Safe to call from several host threads at once (one pool, one queue).
void host_copy_parallel(const CopyJob* jobs, int nJobs);

void set_host_copy_threads(int n);   // total threads working on a copy incl. the caller; 0 = default (X266_HOST_COPY_THREADS or min(8, cpus/2))
int  host_copy_threads();
void set_host_copy_nt(int mask);     // bit 0: non-temporal stores into the pinned slots, bit 1: into caller memory (others: memcpy)
void host_copy_shutdown();           // joins the pool (xGpuFree of the last context does not call it; process exit does)

} // namespace x266
