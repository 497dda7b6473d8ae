// libm_guard.cpp -- run-time guard of the numerics contract.
//
// Parity with the reference is parity with the HOST libm's sincosf, which the reference reaches
// through cexpf (/root/reference/src/complex.c:35).  The device evaluates one specific operation
// sequence: glibc >= 2.28 x86-64 `__sincosf_fma` (sincosf_glibc.h).  On a host whose libm resolves
// sincosf differently (no-FMA ifunc variant, another libc) the GPU output would silently stop
// being "what the reference prints on this box".  This file runs the HOST TWIN of the device routine
// -- the same header compiled for the CPU -- against the host's sincosf on probes spanning all four
// glibc ranges; doppler_b200_create consults it once per process.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/doppler_b200.h"
#include "sincosf_glibc.h"

namespace {

bool same_bits(float a, float b)
{
    if (a != a || b != b) return (a != a) && (b != b);   // NaN: payload / sign not compared
    uint32_t x, y;
    memcpy(&x, &a, 4);
    memcpy(&y, &b, 4);
    return x == y;
}

void host_sincosf(float y, float* s, float* c) { sincosf(y, s, c); }

}  // namespace

extern "C" {

// Number of probes on which `fn` and the twin of the device routine disagree (0 = compatible).
uint32_t doppler_b200_libm_mismatches(void (*fn)(float, float*, float*))
{
    if (!fn) fn = host_sincosf;
    uint32_t bad = 0;
    auto probe = [&](uint32_t bits) {
        float y;
        memcpy(&y, &bits, 4);
        float ls = 0, lc = 0;
        fn(y, &ls, &lc);
        const db_sincos_t r = db_sincosf_glibc(y);
        bad += !(same_bits(ls, r.s) && same_bits(lc, r.c));
    };
    // ranges of the algorithm: tiny < 2^-12, small < pi/4, medium < 120, large (96-bit reduction), non-finite;
    // a multiplicative stride walks every binade of each, both signs; plus the thetas the mixer actually forms
    // for a handful of ratios (theta = -2pi * (r * n), dsp.rs:121)
    uint32_t u = 0x00000001u;
    for (int i = 0; i < 6000; i++) {
        probe(u);
        probe(u | 0x80000000u);
        u += 0x0005a3c7u + (uint32_t)i * 97u;   // ~2^19 per step: ~6000 steps span 0 .. 0x7f800000 and the NaNs behind it
        if (u >= 0x7fc00001u) u = 0x00000001u + (uint32_t)i;
    }
    const float ratios[] = {-15000.0f / 256000.0f, 100000.0f / 10000000.0f, -9876.54f / 1024000.0f, 4000000.5f / 200000000.0f};
    for (float r : ratios)
        for (uint32_t n = 1; n < 4000000u; n += 1999u) {
            const float theta = -6.2831855f * (r * (float)n);
            uint32_t b;
            memcpy(&b, &theta, 4);
            probe(b);
        }
    return bad;
}

int doppler_b200_libm_compatible(void)
{
    static const int ok = doppler_b200_libm_mismatches(nullptr) == 0;
    return ok;
}

}  // extern "C"
