/*
 * tests/c/bdpi_dropin.c -- plain-C proof of the drop-in boundary (no Python, no torch): dlopen()s a library that
 * exports the reference's BDPI symbols and replays the Bluesim testbench call sequence
 * (src/mkDct32.bsv:430-470: genNew, 16 x getDiff, 256 x getDct, 11 blocks; src/mkSatd.bsv:222-252: genNew,
 * 8 x getDiff, getSatd, 256 iterations), printing an FNV-1a-64 of everything it received.  Run once against
 * oracle/_ref/libx266ref.so (the unmodified reference) and once against x266_b200/libx266_b200.so with the same
 * srand() seed: the two digests must be identical.  Also calls the Tier-2 symbols when present.
 *   gcc -O2 tests/c/bdpi_dropin.c -o bdpi_dropin -ldl ;  ./bdpi_dropin <lib.so> <seed>
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

static unsigned long long h = 1469598103934665603ull;
static void mix(const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    size_t i;
    for (i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
}

int main(int argc, char** argv)
{
    void* L;
    void (*dctNew)(void); void (*dctDiff)(unsigned int*); unsigned long long (*dctGet)(void);
    void (*satdNew)(void); void (*satdDiff)(unsigned int*); unsigned int (*satdGet)(void);
    const short (*g)[32];
    int blk, i;
    if (argc < 3) { fprintf(stderr, "usage: %s lib.so seed\n", argv[0]); return 2; }
    L = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!L) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    dctNew = (void (*)(void))dlsym(L, "dct32_genNew");
    dctDiff = (void (*)(unsigned int*))dlsym(L, "dct32_getDiff");
    dctGet = (unsigned long long (*)(void))dlsym(L, "dct32_getDct");
    satdNew = (void (*)(void))dlsym(L, "satd8x8_genNew");
    satdDiff = (void (*)(unsigned int*))dlsym(L, "satd8x8_getDiff");
    satdGet = (unsigned int (*)(void))dlsym(L, "satd8x8_getSatd");
    g = (const short (*)[32])dlsym(L, "g_t32");
    if (!dctNew || !dctDiff || !dctGet || !satdNew || !satdDiff || !satdGet || !g) { fprintf(stderr, "missing BDPI symbol\n"); return 2; }
    mix(g, 32 * 32 * sizeof(short));
    srand((unsigned)atoi(argv[2]));
    for (blk = 0; blk < 11; blk++)
    {
        unsigned int res[32];
        dctNew();
        for (i = 0; i < 16; i++) { dctDiff(res); mix(res, sizeof(res)); }
        for (i = 0; i < 256; i++) { unsigned long long w = dctGet(); mix(&w, sizeof(w)); }
    }
    for (blk = 0; blk < 256; blk++)
    {
        unsigned int res[4], s;
        satdNew();
        for (i = 0; i < 8; i++) { satdDiff(res); mix(res, sizeof(res)); }
        s = satdGet();
        mix(&s, sizeof(s));
    }
    printf("%016llx\n", h);
    return 0;
}
