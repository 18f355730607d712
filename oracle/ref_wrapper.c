/*
 * oracle/ref_wrapper.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Builds the UNMODIFIED reference golden model in place: this translation unit #includes
 * $(X266_REF)/src_tb/dct32.c and $(X266_REF)/src_tb/satd.c where they lie (no reference source is
 * copied into this repository) and re-exports their two `static` kernels under ref_* names, plus a
 * pthread fan-out used as the timed CPU baseline ("cpu_baseline.kind": "reference").
 *
 *  - dct32.c:84-106 has a live debug printf inside the row loop when shift==11; it is compiled out
 *    with a macro (results are unaffected, it only prints).
 *  - both files define `static int16_t mat[]` (dct32.c:173, satd.c:120); renamed per include.
 *  - tb_common.h:29-37 hand-rolls the stdint typedefs (uint64_t = unsigned long long), so this TU
 *    must not include <stdint.h> before it; everything exported uses plain C types.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define printf(...) ((void)0)
#define mat ref_dct_mat
#define dct ref_dct_out
#include "dct32.c"            /* -I$(X266_REF)/src_tb */
#undef mat
#undef dct
#define mat ref_satd_mat
#include "satd.c"
#undef mat
#undef printf

/* --- the two static kernels, exported ------------------------------------------------------- */
void ref_partialButterfly32(const short* src, short* dst, int shift, int line)
{
    partialButterfly32(src, dst, shift, line);     /* dct32.c:66-170 */
}

int ref_satd8x8(const short* diff)
{
    return satd8x8(diff);                          /* satd.c:31-118 */
}

/* 2-D composition exactly as dct32_genNew does it (dct32.c:197-198) but with caller data+shifts */
void ref_dct32_2d(const short* src, short* dst, int shift1, int shift2)
{
    short coef[32 * 32];
    partialButterfly32(src, coef, shift1, 32);
    partialButterfly32(coef, dst, shift2, 32);
}

/* read-back of the module-static state after dct32_genNew()/satd8x8_genNew() */
const short* ref_dct32_lastMat(void) { return ref_dct_mat; }
const short* ref_dct32_lastDct(void) { return ref_dct_out; }
const short* ref_satd_lastMat(void)  { return ref_satd_mat; }

/* --- batch drivers (contiguous ranges per pthread) ------------------------------------------ */
typedef struct { const short* src; short* dst; size_t lo, hi; int s1, s2; } dct_job_t;
typedef struct { const short* src; int* dst; size_t lo, hi; } satd_job_t;

static void* dct_worker(void* p)
{
    dct_job_t* j = (dct_job_t*)p;
    size_t b;
    for (b = j->lo; b < j->hi; b++)
        ref_dct32_2d(j->src + b * 1024, j->dst + b * 1024, j->s1, j->s2);
    return 0;
}

static void* satd_worker(void* p)
{
    satd_job_t* j = (satd_job_t*)p;
    size_t b;
    for (b = j->lo; b < j->hi; b++)
        j->dst[b] = satd8x8(j->src + b * 64);
    return 0;
}

#define REF_MAX_THREADS 1024

int ref_dct32_batch(const short* src, short* dst, size_t n, int s1, int s2, int threads)
{
    pthread_t th[REF_MAX_THREADS];
    dct_job_t jobs[REF_MAX_THREADS];
    int t;
    if (threads < 1) threads = 1;
    if (threads > REF_MAX_THREADS) threads = REF_MAX_THREADS;
    for (t = 0; t < threads; t++)
    {
        jobs[t].src = src; jobs[t].dst = dst; jobs[t].s1 = s1; jobs[t].s2 = s2;
        jobs[t].lo = n * (size_t)t / threads;
        jobs[t].hi = n * (size_t)(t + 1) / threads;
        if (threads == 1) { dct_worker(&jobs[0]); return 0; }
        if (pthread_create(&th[t], 0, dct_worker, &jobs[t])) return -1;
    }
    for (t = 0; t < threads; t++) pthread_join(th[t], 0);
    return 0;
}

int ref_satd8x8_batch(const short* diff, int* out, size_t n, int threads)
{
    pthread_t th[REF_MAX_THREADS];
    satd_job_t jobs[REF_MAX_THREADS];
    int t;
    if (threads < 1) threads = 1;
    if (threads > REF_MAX_THREADS) threads = REF_MAX_THREADS;
    for (t = 0; t < threads; t++)
    {
        jobs[t].src = diff; jobs[t].dst = out;
        jobs[t].lo = n * (size_t)t / threads;
        jobs[t].hi = n * (size_t)(t + 1) / threads;
        if (threads == 1) { satd_worker(&jobs[0]); return 0; }
        if (pthread_create(&th[t], 0, satd_worker, &jobs[t])) return -1;
    }
    for (t = 0; t < threads; t++) pthread_join(th[t], 0);
    return 0;
}
