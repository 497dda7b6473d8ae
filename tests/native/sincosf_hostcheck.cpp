// sincosf_hostcheck.cpp -- TEST-ONLY host twin of the device sincosf.
//
// Compiles doppler_b200/csrc/sincosf_glibc.h for the host and compares it bit-for-bit with
// the host libm's sincosf (the function the reference reaches through src/complex.c:35).
// This is how the device routine's operation sequence is validated without a GPU: IEEE-754
// double add/mul/fma and the integer path are identical on both sides.  Not product code.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../doppler_b200/csrc/sincosf_glibc.h"

extern "C" {

// Single evaluation of the twin (for ctypes spot checks).
void hostcheck_sincosf(float y, float* s, float* c)
{
    db_sincos_t r = db_sincosf_glibc(y);
    *s = r.s;
    *c = r.c;
}

struct sweep_job {
    uint64_t first, count, stride;
    uint64_t mismatches;
    uint32_t first_bad;
};

static void* sweep_run(void* p)
{
    sweep_job* j = (sweep_job*)p;
    j->mismatches = 0;
    j->first_bad = 0;
    for (uint64_t i = 0; i < j->count; i++) {
        uint32_t u = (uint32_t)(j->first + i * j->stride);
        float y;
        memcpy(&y, &u, 4);
        float ls, lc;
        sincosf(y, &ls, &lc);
        db_sincos_t r = db_sincosf_glibc(y);
        uint32_t a, b, c, d;
        memcpy(&a, &ls, 4);
        memcpy(&b, &r.s, 4);
        memcpy(&c, &lc, 4);
        memcpy(&d, &r.c, 4);
        int bad;
        if (ls != ls || lc != lc)
            bad = !(r.s != r.s && r.c != r.c);   // NaN: payload/sign not compared
        else
            bad = (a != b) || (c != d);
        if (bad) {
            if (!j->mismatches) j->first_bad = u;
            j->mismatches++;
        }
    }
    return NULL;
}

// Compare over bit patterns first, first+stride, ... (count of them), split over threads.
// Returns the number of mismatching inputs; *first_bad gets one offending bit pattern.
uint64_t hostcheck_sweep(uint64_t first, uint64_t count, uint64_t stride, int threads, uint32_t* first_bad)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    sweep_job jobs[256];
    pthread_t th[256];
    uint64_t per = (count + threads - 1) / threads, k = 0;
    for (int t = 0; t < threads; t++) {
        uint64_t n = k + per <= count ? per : count - k;
        jobs[t].first = first + k * stride;
        jobs[t].count = n;
        jobs[t].stride = stride;
        k += n;
        pthread_create(&th[t], NULL, sweep_run, &jobs[t]);
    }
    uint64_t bad = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].mismatches && !bad && first_bad) *first_bad = jobs[t].first_bad;
        bad += jobs[t].mismatches;
    }
    return bad;
}


// ---- range-specialised twins (db_sincosf_large / _medium / _small) against libm -------------
// Each bit pattern is evaluated by the routine of ITS glibc range, with the large-range window
// derived from the pattern's own sign and exponent -- exactly what a tile does on the device.
static void* fast_run(void* p)
{
    sweep_job* j = (sweep_job*)p;
    j->mismatches = 0;
    j->first_bad = 0;
    uint32_t cached_key = 0xffffffffu;
    db_window_t win;
    memset(&win, 0, sizeof win);
    for (uint64_t i = 0; i < j->count; i++) {
        const uint32_t u = (uint32_t)(j->first + i * j->stride);
        const uint32_t a = u & 0x7fffffffu;
        float y, fs, fc;
        memcpy(&y, &u, 4);
        if (a >= 0x7f800000u) continue;             // Inf / NaN: generic routine only
        if (a < 0x39800000u) {                      // tiny
            fs = y;
            fc = 1.0f;
        } else if (a < 0x3f400000u) {
            db_sincosf_small(y, &fs, &fc);
        } else if (a < 0x42f00000u) {
            db_sincosf_medium(y, &fs, &fc);
        } else {
            if ((u >> 23) != cached_key) {
                db_large_window(u, &win);
                cached_key = u >> 23;
            }
            db_sincosf_large(u, &win, &fs, &fc);
        }
        float ls, lc;
        sincosf(y, &ls, &lc);
        if (memcmp(&ls, &fs, 4) != 0 || memcmp(&lc, &fc, 4) != 0) {
            if (!j->mismatches) j->first_bad = u;
            j->mismatches++;
        }
    }
    return NULL;
}

uint64_t hostcheck_fast_sweep(uint64_t first, uint64_t count, uint64_t stride, int threads, uint32_t* first_bad)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    sweep_job jobs[256];
    pthread_t th[256];
    uint64_t per = (count + threads - 1) / threads, k = 0;
    for (int t = 0; t < threads; t++) {
        uint64_t n = k + per <= count ? per : count - k;
        jobs[t].first = first + k * stride;
        jobs[t].count = n;
        jobs[t].stride = stride;
        k += n;
        pthread_create(&th[t], NULL, fast_run, &jobs[t]);
    }
    uint64_t bad = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].mismatches && !bad && first_bad) *first_bad = jobs[t].first_bad;
        bad += jobs[t].mismatches;
    }
    return bad;
}

}  // extern "C"

#ifdef HOSTCHECK_MAIN
int main(int argc, char** argv)
{
    uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 0) : 1;
    int threads = argc > 2 ? atoi(argv[2]) : 8;
    uint32_t fb = 0;
    uint64_t count = ((1ULL << 32) + stride - 1) / stride;
    uint64_t bad = hostcheck_sweep(0, count, stride, threads, &fb);
    printf("generic: checked %llu patterns (stride %llu): %llu mismatches", (unsigned long long)count,
           (unsigned long long)stride, (unsigned long long)bad);
    if (bad) printf(" (e.g. 0x%08x)", fb);
    printf("\n");
    uint64_t bad2 = hostcheck_fast_sweep(0, count, stride, threads, &fb);
    printf("range-specialised: checked %llu patterns (stride %llu): %llu mismatches", (unsigned long long)count,
           (unsigned long long)stride, (unsigned long long)bad2);
    if (bad2) printf(" (e.g. 0x%08x)", fb);
    printf("\n");
    return (bad || bad2) ? 1 : 0;
}
#endif
