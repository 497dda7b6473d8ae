// sincosf_glibc.h -- bit-exact re-implementation of glibc 2.39 x86-64 `sincosf`
// (the `__sincosf_fma` ifunc variant) for CUDA device code, with a host twin for testing.
//
// Why: the reference's oscillator is libm `cexpf(0 + i*theta)` called once per sample
// (/root/reference/src/complex.c:33-39, call site src/dsp.rs:121-122).  For a zero real part
// glibc's cexpf reduces to exp(0) * sincosf(theta) = (cos, sin) exactly, so parity with the
// reference is parity with the host libm's sincosf -- to the last bit, because the i16 egress
// cast truncates (src/main.rs:77-78) and a 1-ULP trig difference flips output samples.
//
// glibc's source is not in this image; the operation sequence below was recovered from the
// disassembly of /lib/x86_64-linux-gnu/libm.so.6 (GLIBC 2.39-0ubuntu8.5, function at 0x7e570,
// selected by the sincosf ifunc when the CPU has FMA+AVX2) and the constants were read from
// its .rodata (0xb80c0 inv_pio4[24], 0xb8120 sincos table, 0x99ec0 pi63).  The algorithm is
// the published ARM "optimized routines" sincosf (glibc >= 2.28): evaluate in double, three
// ranges (|y| < pi/4; |y| < 120: one-step reduction by pi/2; otherwise a 96-bit fixed-point
// 4/pi multiply), one degree-7/8 polynomial pair, round to float once.  In the FMA build
// every `a + b*c` of the C source is ONE fused operation; this file spells each fused /
// unfused operation explicitly so no compiler flag can change it.
//
// Verified: tests/native/sincosf_hostcheck.cpp (driven by tests/test_sincosf_host.py) compares
// db_sincosf_glibc() against the host's libm sincosf over all 2^32 float bit patterns (bit-identical,
// NaN payloads aside; DOPPLER_FULL_SWEEP=1, strided by default);
// tests/test_gpu_parity.py::test_device_sincosf_matches_host_libm does the same on the device, and
// libm_guard.cpp re-checks a sample of it at run time (doppler_b200_libm_compatible).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DB_HD __host__ __device__ __forceinline__
#else
#define DB_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define DB_DMUL(a, b) __dmul_rn((a), (b))
#define DB_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define DB_F2D(a) ((double)(a))
#define DB_D2F(a) __double2float_rn(a)
#define DB_D2I_RZ(a) __double2int_rz(a)
#define DB_I2D(a) __int2double_rn(a)
#define DB_LL2D(a) __ll2double_rn(a)
#define DB_FBITS(f) __float_as_uint(f)
#define DB_UBITS(u) __uint_as_float(u)
#else
#include <math.h>
#include <string.h>
#define DB_DMUL(a, b) ((a) * (b))
#define DB_DFMA(a, b, c) __builtin_fma((a), (b), (c))
#define DB_F2D(a) ((double)(a))
#define DB_D2F(a) ((float)(a))
#define DB_D2I_RZ(a) ((int32_t)(a))
#define DB_I2D(a) ((double)(a))
#define DB_LL2D(a) ((double)(a))
static inline uint32_t db_fbits_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float db_ubits_host(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define DB_FBITS(f) db_fbits_host(f)
#define DB_UBITS(u) db_ubits_host(u)
#endif

// 4/pi as overlapping 32-bit windows, one per 8 bits (libm .rodata 0xb80c0).
#define DB_INV_PIO4_WORDS                                                                        \
    0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,  \
    0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu,   \
    0xf534ddc0u, 0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u,   \
    0x993c4390u, 0x3c439041u

#if defined(__CUDACC__)
__device__ __constant__ uint32_t db_inv_pio4_dev[24] = {DB_INV_PIO4_WORDS};
#endif
#if !defined(__CUDA_ARCH__)
static const uint32_t db_inv_pio4_host[24] = {DB_INV_PIO4_WORDS};
#endif

struct db_sincos_t {
    float s;
    float c;
};

// Polynomial pair on the reduced argument (libm .rodata 0xb8120: c0 c1 s1 c2 s2 c3 s3 c4).
// Returns sin-poly and cos-poly of xr, both rounded to float.  Odd/even symmetry of every
// operation makes sign application after rounding exact, so callers fold quadrant signs in
// as float sign flips instead of multiplying x by +-1 and switching coefficient tables.
DB_HD void db_sincosf_poly(double xr, float* sp, float* cp)
{
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                 C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    double x2 = DB_DMUL(xr, xr);
    double x3 = DB_DMUL(x2, xr);
    double x4 = DB_DMUL(x2, x2);
    double c1 = DB_DFMA(x2, C1, C0);   // c0 + x2*c1   (fused)
    double s1 = DB_DFMA(x2, S3, S2);   // s2 + x2*s3   (fused)
    double c2 = DB_DFMA(x2, C4, C3);   // c3 + x2*c4   (fused)
    double x5 = DB_DMUL(x2, x3);
    double x6 = DB_DMUL(x2, x4);
    double s = DB_DFMA(x3, S1, xr);    // x + x3*s1    (fused)
    double c = DB_DFMA(x4, C2, c1);    // c1 + x4*c2   (fused)
    *sp = DB_D2F(DB_DFMA(s1, x5, s));  // s + x5*s1    (fused), then round to float
    *cp = DB_D2F(DB_DFMA(c2, x6, c));  // c + x6*c2    (fused), then round to float
}

// sincosf(y) exactly as the host libm computes it.  NaN/Inf -> NaN for both.
DB_HD db_sincos_t db_sincosf_glibc(float y)
{
    db_sincos_t out;
    const uint32_t xi = DB_FBITS(y);
    const uint32_t top12 = (xi >> 20) & 0x7ffu;
    float sp, cp;
    uint32_t n;   // quadrant used for the sin/cos swap
    uint32_t q;   // quadrant used for the signs
    if (top12 < 0x3f4u) {               // |y| < pi/4
        if (top12 < 0x398u) {           // |y| < 2^-12: sin = y, cos = 1
            out.s = y;
            out.c = 1.0f;
            return out;
        }
        db_sincosf_poly(DB_F2D(y), &sp, &cp);
        out.s = sp;
        out.c = cp;
        return out;
    } else if (top12 < 0x42fu) {        // |y| < 120: reduce_fast, non-TOINT_INTRINSICS form
        const double HPI_INV_2P24 = 0x1.45f306dc9c883p+23;   // 2/pi * 2^24
        const double HPI = 0x1.921fb54442d18p+0;
        double x = DB_F2D(y);
        double r = DB_DMUL(x, HPI_INV_2P24);
        int32_t ni = (DB_D2I_RZ(r) + 0x800000) >> 24;
        double xr = DB_DFMA(-DB_I2D(ni), HPI, x);            // x - n*hpi   (fused)
        db_sincosf_poly(xr, &sp, &cp);
        n = (uint32_t)ni;
        q = n;
    } else if (top12 < 0x7f8u) {        // finite: reduce_large
#if defined(__CUDA_ARCH__)
        const uint32_t* arr = &db_inv_pio4_dev[(xi >> 26) & 15u];
#else
        const uint32_t* arr = &db_inv_pio4_host[(xi >> 26) & 15u];
#endif
        const uint32_t shift = (xi >> 23) & 7u;
        uint32_t m = ((xi & 0x7fffffu) | 0x800000u) << shift;
        uint64_t res0 = (uint64_t)(uint32_t)(m * arr[0]);
        uint64_t res1 = (uint64_t)m * arr[4];
        uint64_t res2 = (uint64_t)m * arr[8];
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        uint64_t nn = (res0 + (1ULL << 61)) >> 62;
        res0 -= nn << 62;
        const double PI63 = 0x1.921fb54442d18p-62;
        double xr = DB_DMUL(DB_LL2D((int64_t)res0), PI63);
        db_sincosf_poly(xr, &sp, &cp);
        n = (uint32_t)nn;
        q = n + (xi >> 31);
    } else {                            // Inf / NaN: y - y
        out.s = out.c = y - y;
        return out;
    }
    // sign[q&3] = {+,-,-,+} on the sine polynomial; (q & 2) negates the cosine polynomial;
    // (n & 1) swaps which output each polynomial lands in.
    uint32_t sbits = DB_FBITS(sp) ^ (((q + 1u) & 2u) << 30);
    uint32_t cbits = DB_FBITS(cp) ^ ((q & 2u) << 30);
    if (n & 1u) {
        out.s = DB_UBITS(cbits);
        out.c = DB_UBITS(sbits);
    } else {
        out.s = DB_UBITS(sbits);
        out.c = DB_UBITS(cbits);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Range-specialised evaluation: the same operation sequences as db_sincosf_glibc(), split by
// glibc range for callers that already know (warp-uniformly) which range every argument of a
// tile falls in (mixer_kernels.cuh, "Direct evaluation, fast rows").  No range branches, no
// per-argument 4/pi window lookup, and on the device the double constants are read from the
// constant bank as instruction operands instead of being re-materialised per use.
//   db_sincosf_large : |y| >= 120, finite, all arguments of the tile in ONE binade.  glibc's
//       reduce_large multiplies (m << shift) by a 96-bit window of 4/pi selected by the exponent;
//       both depend only on the exponent, so the window is shifted once per tile
//       (W = window << (e & 7), mod 2^96) and m * W == (m << shift) * window (mod 2^96), which is
//       all that res0 = floor(product / 2^32) mod 2^64 keeps.
//   db_sincosf_medium: 0.75 <= |y| < 120 (reduce_fast).
//   db_sincosf_small : 2^-12 <= |y| < 0.75 (polynomial on y itself).
// The host twins are checked against libm over every float of each range (tests/native).
#define DB_KC_VALUES                                                                               \
    0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10,                   \
    0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13,   \
    0x1.921fb54442d18p-62, 0x1.45f306dc9c883p+23, 0x1.921fb54442d18p+0
// index:  0 C0  1 C1  2 C2  3 C3  4 C4  5 S1  6 S2  7 S3  8 pi*2^-63  9 2/pi*2^24  10 pi/2
#if defined(__CUDACC__)
__device__ __constant__ double db_kc_dev[11] = {DB_KC_VALUES};
#endif
#if !defined(__CUDA_ARCH__)
static const double db_kc_host[11] = {DB_KC_VALUES};
#endif
#if defined(__CUDA_ARCH__)
#define DB_K(i) db_kc_dev[i]
#else
#define DB_K(i) db_kc_host[i]
#endif

DB_HD void db_poly_fast(double xr, float* sp, float* cp)
{
    const double x2 = DB_DMUL(xr, xr);
    const double x3 = DB_DMUL(x2, xr);
    const double x4 = DB_DMUL(x2, x2);
    const double c1 = DB_DFMA(x2, DB_K(1), DB_K(0));
    const double s1 = DB_DFMA(x2, DB_K(7), DB_K(6));
    const double c2 = DB_DFMA(x2, DB_K(4), DB_K(3));
    const double x5 = DB_DMUL(x2, x3);
    const double x6 = DB_DMUL(x2, x4);
    const double s = DB_DFMA(x3, DB_K(5), xr);
    const double c = DB_DFMA(x4, DB_K(2), c1);
    *sp = DB_D2F(DB_DFMA(s1, x5, s));
    *cp = DB_D2F(DB_DFMA(c2, x6, c));
}

struct db_window_t {
    uint32_t w0, w1, w2;   // (4/pi window << (e & 7)) mod 2^96, most significant word first
    uint32_t ks, kc;       // (sign + 1) << 30 and sign << 30, sign = (y < 0)
};

// Window and sign constants for every y that shares theta_bits' sign and exponent.
DB_HD void db_large_window(uint32_t theta_bits, db_window_t* t)
{
    const uint32_t e = (theta_bits >> 23) & 0xffu, idx = (e >> 3) & 15u, sh = e & 7u;
#if defined(__CUDA_ARCH__)
    const uint32_t* arr = &db_inv_pio4_dev[idx];
#else
    const uint32_t* arr = &db_inv_pio4_host[idx];
#endif
    const uint32_t q0 = arr[0], q1 = arr[4], q2 = arr[8];
    t->w0 = sh ? ((q0 << sh) | (q1 >> (32u - sh))) : q0;
    t->w1 = sh ? ((q1 << sh) | (q2 >> (32u - sh))) : q1;
    t->w2 = q2 << sh;
    const uint32_t sign = theta_bits >> 31;
    t->ks = (sign + 1u) << 30;
    t->kc = sign << 30;
}

DB_HD void db_sincosf_large(uint32_t xi, const db_window_t* t, float* s, float* c)
{
#if defined(__CUDA_ARCH__)
    uint32_t m;
    asm("lop3.b32 %0, %1, 0x7fffff, 0x800000, 0xEA;" : "=r"(m) : "r"(xi));   // (xi & 0x7fffff) | 0x800000
#else
    const uint32_t m = (xi & 0x7fffffu) | 0x800000u;
#endif
    const uint32_t lo0 = m * t->w0;
    const uint64_t p2 = (uint64_t)m * t->w2;
    const uint64_t acc = (uint64_t)m * t->w1 + (((uint64_t)lo0 << 32) | (p2 >> 32));   // res0
    const uint32_t hi = (uint32_t)(acc >> 32), lo = (uint32_t)acc;
    const uint32_t q = hi + 0x20000000u;            // n = q >> 30
    const uint32_t hi2 = hi - (q & 0xc0000000u);    // res0 - (n << 62)
    const double xr = DB_DMUL(DB_LL2D((int64_t)(((uint64_t)hi2 << 32) | lo)), DB_K(8));
    float sp, cp;
    db_poly_fast(xr, &sp, &cp);
    const uint32_t sb = DB_FBITS(sp) ^ ((q + t->ks) & 0x80000000u);   // (n + sign + 1) & 2
    const uint32_t cb = DB_FBITS(cp) ^ ((q + t->kc) & 0x80000000u);   // (n + sign) & 2
    const bool swap = (q & 0x40000000u) != 0;                         // n & 1
    *s = DB_UBITS(swap ? cb : sb);
    *c = DB_UBITS(swap ? sb : cb);
}

DB_HD void db_sincosf_medium(float y, float* s, float* c)
{
    const double x = DB_F2D(y);
    const int32_t v = DB_D2I_RZ(DB_DMUL(x, DB_K(9))) + 0x800000;   // n = v >> 24 (arithmetic)
    const double xr = DB_DFMA(-DB_I2D(v >> 24), DB_K(10), x);
    float sp, cp;
    db_poly_fast(xr, &sp, &cp);
    const uint32_t sb = DB_FBITS(sp) ^ ((((uint32_t)v + 0x1000000u) << 6) & 0x80000000u);   // (n + 1) & 2
    const uint32_t cb = DB_FBITS(cp) ^ (((uint32_t)v << 6) & 0x80000000u);                  // n & 2
    const bool swap = ((uint32_t)v & 0x1000000u) != 0;                                      // n & 1
    *s = DB_UBITS(swap ? cb : sb);
    *c = DB_UBITS(swap ? sb : cb);
}

DB_HD void db_sincosf_small(float y, float* s, float* c) { db_poly_fast(DB_F2D(y), s, c); }
