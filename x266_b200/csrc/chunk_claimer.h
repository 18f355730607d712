// chunk_claimer.h -- where the next chunk of a host-pointer call comes from (plain C++, no CUDA: compiled by g++ in tests/c/hostcopy_test.cpp).
#pragma once
#include <stddef.h>
#include <atomic>

namespace x266 {

// Where the next chunk of a call comes from.  A single-GPU call walks its range in order; the multi-GPU entry point shares ONE
// claimer between its per-device threads, so a GPU with a slower host link simply claims fewer chunks (the links of one box are
// not equal: profiles/r02_link_ceiling.md).
struct ChunkClaimer {
    std::atomic<size_t> next{0};
    size_t nUnits = 0, chunk = 0;
    bool shared = false;               // several pipelines claim from this counter: each may only claim when one of its slots is free
    ChunkClaimer(size_t n, size_t c, bool sh = false) : nUnits(n), chunk(c), shared(sh) {}
    bool claim(size_t* u0, size_t* nu)
    {
        const size_t at = next.fetch_add(chunk, std::memory_order_relaxed);
        if (at >= nUnits) return false;
        *u0 = at;
        *nu = (nUnits - at) < chunk ? (nUnits - at) : chunk;
        return true;
    }
};

} // namespace x266
