/*
 * doppler_oracle.c -- CPU ORACLE for the doppler NCO-mixer hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (doppler_b200/csrc, libdoppler_b200.so, bin/doppler) never links or calls it.
 *
 * It is a plain-C restatement of the reference algorithm (cubehub/doppler @ 5f13df14,
 * crate v1.1.10).  The reference is Rust and cannot be built in this image (no rustc /
 * cargo), except for src/complex.c which oracle/Makefile compiles UNMODIFIED from
 * /root/reference into oracle/_ref/libcomplex_ref.so.  By default the oracle calls its own
 * restatement of ccexpf(); oracle_set_ccexpf() swaps in the reference's compiled one so
 * that the arithmetic core is the reference's own object code (tests do both and demand
 * bit-identical results).
 *
 * Pinned against: the reference's only known-answer test, test_cexpf (src/dsp.rs:57-83);
 * see tests/test_oracle.py.  The reference holds NO golden vector for the mixer loop,
 * converters, samplenum rule or egress casts, so those are pinned by
 * executing this restatement linked against the reference's complex.c.
 *
 * Third-party arithmetic not under /root/reference:
 *   - glibc libm cexpf -> sincosf (whatever the host has; this image: glibc 2.39, x86-64,
 *     ifunc selects __sincosf_fma).  Call site: src/complex.c:35.
 *   - num-complex 0.1.35 `Mul` (Cargo.lock:132-139): (a*c - b*d, a*d + b*c), unfused.
 *   - Rust `as` casts: f32 -> i16 saturating with NaN -> 0 (Rust >= 1.45 semantics).
 *   - Rust `u32 += 1` in release mode wraps.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off -fno-fast-math: no FMA contraction,
 * matching rustc which never contracts).
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <float.h>

#if FLT_EVAL_METHOD != 0
#error "oracle needs float expressions evaluated in float (SSE2), as rustc does"
#endif
#ifdef __FAST_MATH__
#error "oracle must not be built with -ffast-math"
#endif

typedef struct {
    float re;
    float im;
} oracle_c32;

enum { ORACLE_I16 = 0, ORACLE_F32 = 1 };

/* src/main.rs:49 */
#define ORACLE_BUFFER_SIZE 8192

/* ------------------------------------------------------------------------------------ */
/* src/complex.c:33-39  ccexpf: in-place z <- cexpf(z) through a {float,float} struct.  */
static void oracle_ccexpf_restated(oracle_c32* a)
{
    float complex input = a->re + a->im * I;
    float complex out = cexpf(input);
    a->re = crealf(out);
    a->im = cimagf(out);
}

typedef void (*oracle_ccexpf_fn)(oracle_c32*);
static oracle_ccexpf_fn g_ccexpf = oracle_ccexpf_restated;

/* Swap in the reference's own compiled ccexpf (oracle/_ref/libcomplex_ref.so); NULL restores
 * the restatement. */
void oracle_set_ccexpf(oracle_ccexpf_fn fn) { g_ccexpf = fn ? fn : oracle_ccexpf_restated; }

void oracle_ccexpf(oracle_c32* z) { g_ccexpf(z); }

/* ------------------------------------------------------------------------------------ */
/* src/dsp.rs:85-99  convert_iqi16_to_complex.  Returns sample count, or -1 where the     */
/* reference's assert!(len % 4 == 0) (dsp.rs:87) would panic.                             */
long oracle_convert_iqi16_to_complex(const uint8_t* inbuf, size_t len, oracle_c32* out)
{
    if (len % 4 != 0) return -1;
    size_t n = 0;
    for (size_t p = 0; p < len; p += 4) {
        const uint8_t* b = inbuf + p;
        int16_t iv = (int16_t)(uint16_t)(((uint16_t)b[1] << 8) | (uint16_t)b[0]);
        int16_t qv = (int16_t)(uint16_t)(((uint16_t)b[3] << 8) | (uint16_t)b[2]);
        out[n].re = (float)iv / 32768.0f;
        out[n].im = (float)qv / 32768.0f;
        n++;
    }
    return (long)n;
}

/* src/dsp.rs:101-115  convert_iqf32_to_complex: LE byte reassembly + transmute (bit copy).
 * Returns -1 where assert!(len % 8 == 0) (dsp.rs:103) would panic. */
long oracle_convert_iqf32_to_complex(const uint8_t* inbuf, size_t len, oracle_c32* out)
{
    if (len % 8 != 0) return -1;
    size_t n = 0;
    for (size_t p = 0; p < len; p += 8) {
        const uint8_t* b = inbuf + p;
        uint32_t iu = ((uint32_t)b[3] << 24) | ((uint32_t)b[2] << 16) | ((uint32_t)b[1] << 8) | b[0];
        uint32_t qu = ((uint32_t)b[7] << 24) | ((uint32_t)b[6] << 16) | ((uint32_t)b[5] << 8) | b[4];
        memcpy(&out[n].re, &iu, 4);
        memcpy(&out[n].im, &qu, 4);
        n++;
    }
    return (long)n;
}

/* ------------------------------------------------------------------------------------ */
/* src/dsp.rs:125  (shift_hz / samplerate as f32 * *samplenum as f32).fract() == 0.0      */
/* f32::fract(x) = x - x.trunc(); inf -> NaN -> compares false.                           */
static inline int oracle_reset_test(float shift_hz, uint32_t samplerate, uint32_t samplenum)
{
    float x = shift_hz / (float)samplerate * (float)samplenum;
    float fr = x - truncf(x);
    return fr == 0.0f;
}

/* src/dsp.rs:117-134  shift_frequency.  `samplenum` is the in/out u32 of src/main.rs:60. */
void oracle_shift_frequency(const oracle_c32* inbuf, size_t n, uint32_t* samplenum, float shift_hz,
                            uint32_t samplerate, oracle_c32* out)
{
    const float PI_F32 = 3.14159274101257324219f; /* std::f32::consts::PI */
    for (size_t k = 0; k < n; k++) {
        /* dsp.rs:121: Complex::new(0.0, -2. * PI * (shift_hz / samplerate as f32 * n as f32)) */
        float x = shift_hz / (float)samplerate * (float)(*samplenum);
        float c = -2.0f * PI_F32;
        oracle_c32 corrector;
        corrector.re = 0.0f;
        corrector.im = c * x;
        g_ccexpf(&corrector); /* dsp.rs:122 */

        /* dsp.rs:123  sample * corrector  (num-complex Mul: re = a*c - b*d, im = a*d + b*c) */
        float a = inbuf[k].re, b = inbuf[k].im;
        float ac = a * corrector.re;
        float bd = b * corrector.im;
        float ad = a * corrector.im;
        float bc = b * corrector.re;
        out[k].re = ac - bd;
        out[k].im = ad + bc;

        /* dsp.rs:125-130 */
        if (oracle_reset_test(shift_hz, samplerate, *samplenum))
            *samplenum = 1;
        else
            *samplenum += 1; /* release-mode wrap */
    }
}

/* Only the samplenum recurrence of dsp.rs:125-130 (no trig): the reference state after
 * `count` samples at constant shift.  Used to seed contiguous slices in the threaded baseline
 * and as the checker for the product's analytic planner. */
uint32_t oracle_samplenum_advance(uint32_t samplenum, float shift_hz, uint32_t samplerate, uint64_t count)
{
    for (uint64_t k = 0; k < count; k++) {
        if (oracle_reset_test(shift_hz, samplerate, samplenum))
            samplenum = 1;
        else
            samplenum += 1;
    }
    return samplenum;
}

/* Writes the per-sample samplenum sequence (value USED for sample k) -- checker for the
 * product planner's piece table. */
uint32_t oracle_samplenum_trace(uint32_t samplenum, float shift_hz, uint32_t samplerate, uint64_t count,
                                uint32_t* trace)
{
    for (uint64_t k = 0; k < count; k++) {
        trace[k] = samplenum;
        if (oracle_reset_test(shift_hz, samplerate, samplenum))
            samplenum = 1;
        else
            samplenum += 1;
    }
    return samplenum;
}

/* ------------------------------------------------------------------------------------ */
/* Rust `f32 as i16`: truncate toward zero, saturate, NaN -> 0. */
static inline int16_t oracle_f32_as_i16(float v)
{
    if (v != v) return 0;
    if (v >= 32767.0f) return 32767;
    if (v <= -32768.0f) return -32768;
    return (int16_t)v; /* C truncates toward zero; in range here */
}

/* src/main.rs:73-87  i16 egress: (re * 32767.0) as i16, little-endian bytes. */
size_t oracle_egress_i16(const oracle_c32* in, size_t n, uint8_t* out)
{
    for (size_t k = 0; k < n; k++) {
        float fi = in[k].re * 32767.0f;
        float fq = in[k].im * 32767.0f;
        int16_t i = oracle_f32_as_i16(fi);
        int16_t q = oracle_f32_as_i16(fq);
        out[4 * k + 0] = (uint8_t)(i & 0xFF);
        out[4 * k + 1] = (uint8_t)((i >> 8) & 0xFF);
        out[4 * k + 2] = (uint8_t)(q & 0xFF);
        out[4 * k + 3] = (uint8_t)((q >> 8) & 0xFF);
    }
    return 4 * n;
}

/* src/main.rs:89-93  f32 egress: raw native-endian bytes of Vec<Complex<f32>>. */
size_t oracle_egress_f32(const oracle_c32* in, size_t n, uint8_t* out)
{
    memcpy(out, in, 8 * n);
    return 8 * n;
}

/* ------------------------------------------------------------------------------------ */
/* src/main.rs:62-99  closure `shift`: one 8192-byte block: convert -> shift_frequency ->  */
/* egress.  `avail` is what the `take(BUFFER_SIZE)` read returned.  Returns samples        */
/* processed, or -1 for the converters' assert panic.  *out_len receives bytes written.    */
static long oracle_block(const uint8_t* in, size_t avail, int intype, int outtype, float shift_hz,
                         uint32_t samplerate, uint32_t* samplenum, uint8_t* out, size_t* out_len,
                         oracle_c32* scratch_in, oracle_c32* scratch_out)
{
    long n = (intype == ORACLE_I16) ? oracle_convert_iqi16_to_complex(in, avail, scratch_in)
                                    : oracle_convert_iqf32_to_complex(in, avail, scratch_in);
    if (n < 0) return -1;
    oracle_shift_frequency(scratch_in, (size_t)n, samplenum, shift_hz, samplerate, scratch_out);
    *out_len = (outtype == ORACLE_I16) ? oracle_egress_i16(scratch_out, (size_t)n, out)
                                       : oracle_egress_f32(scratch_out, (size_t)n, out);
    return n;
}

/* Fused single call on an arbitrary-length buffer with ONE shift value: what a caller of the
 * library boundary (convert -> shift_frequency -> egress, main.rs:65-94) computes when it
 * passes the whole buffer at once.  Returns bytes written or -1. */
long oracle_mix(const uint8_t* in, size_t len, int intype, int outtype, float shift_hz, uint32_t samplerate,
                uint32_t* samplenum, uint8_t* out)
{
    size_t bps = (intype == ORACLE_I16) ? 4 : 8;
    if (len % bps != 0) return -1;
    size_t n = len / bps;
    oracle_c32* a = (oracle_c32*)malloc((n ? n : 1) * sizeof(oracle_c32));
    oracle_c32* b = (oracle_c32*)malloc((n ? n : 1) * sizeof(oracle_c32));
    size_t out_len = 0;
    long r = oracle_block(in, len, intype, outtype, shift_hz, samplerate, samplenum, out, &out_len, a, b);
    free(a);
    free(b);
    return r < 0 ? -1 : (long)out_len;
}

/* src/main.rs:102-119  const-mode driver over an in-memory "stdin".  shift is the CLI's i32
 * (main.rs:110 `as f32`).  Stops on the first short read (main.rs:98,115).  Returns bytes
 * written, or -(bytes written before the panic) - 1 if a converter assert would fire. */
long oracle_const_stream(const uint8_t* in, size_t len, int intype, int outtype, int32_t shift, uint32_t samplerate,
                         uint8_t* out, uint32_t* samplenum_out)
{
    oracle_c32 a[ORACLE_BUFFER_SIZE / 4], b[ORACLE_BUFFER_SIZE / 4];
    uint32_t samplenr = 0; /* main.rs:60 */
    float shift_hz = (float)shift;
    size_t pos = 0, wr = 0;
    for (;;) {
        size_t avail = len - pos < ORACLE_BUFFER_SIZE ? len - pos : ORACLE_BUFFER_SIZE;
        size_t out_len = 0;
        long n = oracle_block(in + pos, avail, intype, outtype, shift_hz, samplerate, &samplenr, out + wr, &out_len, a, b);
        if (n < 0) return -(long)wr - 1;
        pos += avail;
        wr += out_len;
        if (avail != ORACLE_BUFFER_SIZE) break;
    }
    if (samplenum_out) *samplenum_out = samplenr;
    return (long)wr;
}

/* src/main.rs:163  doppler_hz from the propagator's range rate (f64 arithmetic, as written:
 * ((rr * 1000 / c) * f as f64) * (-1.0)). */
double oracle_doppler_hz(double range_rate_km_sec, uint32_t frequency)
{
    const double SPEED_OF_LIGHT_M_S = 299792458.0; /* main.rs:48 */
    return (range_rate_km_sec * 1000.0 / SPEED_OF_LIGHT_M_S) * (double)frequency * (-1.0);
}

/* src/main.rs:155-184  track-mode replay driver (--time given).  The propagator
 * (predict.update(start + dt); libgpredict, not in /root/reference) is abstracted as a table:
 * doppler_hz_by_second[s] = the f64 doppler_hz that main.rs:163 yields at start_time + s
 * seconds (index clamped to nsec-1).  Everything else -- one-block lag of dt, f32 time
 * arithmetic (main.rs:166), `doppler_hz as f32 + offset as f32` (main.rs:177), samplenum
 * carried across shift changes (main.rs:60) -- follows the reference.  If shifts_out != NULL
 * it receives the f32 shift used for every block (including the final short one), capacity
 * shifts_cap; *nblocks_out = number of blocks pumped. */
long oracle_track_replay_stream(const uint8_t* in, size_t len, int intype, int outtype, const double* doppler_hz_by_second,
                                size_t nsec, int32_t offset, uint32_t samplerate, uint8_t* out, uint32_t* samplenum_out,
                                float* shifts_out, size_t shifts_cap, size_t* nblocks_out)
{
    oracle_c32 a[ORACLE_BUFFER_SIZE / 4], b[ORACLE_BUFFER_SIZE / 4];
    uint32_t samplenr = 0;
    size_t sample_count = 0; /* main.rs:157 */
    int64_t dt = 0;          /* main.rs:158, whole seconds */
    size_t pos = 0, wr = 0, nb = 0;
    for (;;) {
        size_t idx = (size_t)dt < nsec ? (size_t)dt : nsec - 1;
        double doppler_hz = doppler_hz_by_second[idx]; /* main.rs:162-163 at start + dt */
        dt = (int64_t)((float)sample_count / (float)samplerate); /* main.rs:166 */
        float shift_hz = (float)doppler_hz + (float)offset;     /* main.rs:177 */
        if (shifts_out && nb < shifts_cap) shifts_out[nb] = shift_hz;
        nb++;
        size_t avail = len - pos < ORACLE_BUFFER_SIZE ? len - pos : ORACLE_BUFFER_SIZE;
        size_t out_len = 0;
        long n = oracle_block(in + pos, avail, intype, outtype, shift_hz, samplerate, &samplenr, out + wr, &out_len, a, b);
        if (n < 0) return -(long)wr - 1;
        pos += avail;
        wr += out_len;
        if (avail != ORACLE_BUFFER_SIZE) break; /* main.rs:178-180 */
        sample_count += (size_t)n;              /* main.rs:182 */
    }
    if (samplenum_out) *samplenum_out = samplenr;
    if (nblocks_out) *nblocks_out = nb;
    return (long)wr;
}

/* Per-block shift array variant (the planned entry point's checker): block b of
 * ORACLE_BUFFER_SIZE input bytes uses shifts[b]; samplenum carried. */
long oracle_mix_blocks(const uint8_t* in, size_t len, int intype, int outtype, const float* shifts, size_t nshifts,
                       uint32_t samplerate, uint32_t* samplenum, uint8_t* out)
{
    oracle_c32 a[ORACLE_BUFFER_SIZE / 4], b[ORACLE_BUFFER_SIZE / 4];
    size_t pos = 0, wr = 0, nb = 0;
    while (pos < len) {
        if (nb >= nshifts) return -(long)wr - 1;
        size_t avail = len - pos < ORACLE_BUFFER_SIZE ? len - pos : ORACLE_BUFFER_SIZE;
        size_t out_len = 0;
        long n = oracle_block(in + pos, avail, intype, outtype, shifts[nb], samplerate, samplenum, out + wr, &out_len, a, b);
        if (n < 0) return -(long)wr - 1;
        pos += avail;
        wr += out_len;
        nb++;
    }
    return (long)wr;
}

/* ------------------------------------------------------------------------------------ */
/* CPU baseline timing (bench.py cpu_baseline / --impl reference): the same restated loop   */
/* over an in-memory buffer, `threads` contiguous slices, each seeded with the reference     */
/* samplenum at its first sample (state chained by oracle_samplenum_advance before timing).  */
typedef struct {
    const uint8_t* in;
    uint8_t* out;
    size_t nsamples;
    int intype, outtype;
    float shift_hz;
    uint32_t samplerate;
    uint32_t samplenum;
} oracle_slice;

static void* oracle_slice_run(void* p)
{
    oracle_slice* s = (oracle_slice*)p;
    oracle_c32 a[ORACLE_BUFFER_SIZE / 4], b[ORACLE_BUFFER_SIZE / 4];
    size_t ibps = s->intype == ORACLE_I16 ? 4 : 8, obps = s->outtype == ORACLE_I16 ? 4 : 8;
    size_t done = 0;
    uint32_t sn = s->samplenum;   /* thread-local: the per-sample state update must not share a cache line with other threads' */
    while (done < s->nsamples) {
        /* the reference's own block size in samples (8192 bytes, main.rs:49) */
        size_t n = s->nsamples - done;
        size_t blk = ORACLE_BUFFER_SIZE / ibps;
        if (n > blk) n = blk;
        size_t out_len = 0;
        oracle_block(s->in + done * ibps, n * ibps, s->intype, s->outtype, s->shift_hz, s->samplerate, &sn,
                     s->out + done * obps, &out_len, a, b);
        done += n;
    }
    s->samplenum = sn;
    return NULL;
}

/* Returns elapsed seconds of the threaded mixing (state seeding excluded), <0 on error. */
double oracle_bench_const(const uint8_t* in, size_t nsamples, int intype, int outtype, float shift_hz, uint32_t samplerate,
                          uint8_t* out, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    oracle_slice* sl = (oracle_slice*)calloc((size_t)threads, sizeof(oracle_slice));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    size_t ibps = intype == ORACLE_I16 ? 4 : 8, obps = outtype == ORACLE_I16 ? 4 : 8;
    size_t per = (nsamples + (size_t)threads - 1) / (size_t)threads;
    uint32_t sn = 0;
    size_t k = 0;
    for (int t = 0; t < threads; t++) {
        size_t n = k + per <= nsamples ? per : nsamples - k;
        sl[t] = (oracle_slice){in + k * ibps, out + k * obps, n, intype, outtype, shift_hz, samplerate, sn};
        sn = oracle_samplenum_advance(sn, shift_hz, samplerate, n);
        k += n;
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, oracle_slice_run, &sl[t]);
    oracle_slice_run(&sl[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(sl);
    free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* The same for a per-block shift schedule (track mode): `threads` contiguous runs of whole 8192-byte
 * blocks, thread t seeded with the samplenum the sequential recurrence (dsp.rs:125-130) reaches at its
 * first sample -- computed here by running that recurrence, not by the product's analytic planner, so
 * the output is independent of the code under test.  Returns the elapsed seconds of the threaded mixing
 * (seeding excluded); *samplenum is advanced to the state after the last sample. */
typedef struct {
    const uint8_t* in;
    uint8_t* out;
    size_t len;          /* bytes */
    int intype, outtype;
    const float* shifts;
    size_t nshifts;
    uint32_t samplerate;
    uint32_t samplenum;
    long rc;
} oracle_blocks_slice;

static void* oracle_blocks_run(void* p)
{
    oracle_blocks_slice* s = (oracle_blocks_slice*)p;
    uint32_t sn = s->samplenum;   /* thread-local, see oracle_slice_run */
    s->rc = oracle_mix_blocks(s->in, s->len, s->intype, s->outtype, s->shifts, s->nshifts, s->samplerate, &sn, s->out);
    s->samplenum = sn;
    return NULL;
}

double oracle_bench_blocks(const uint8_t* in, size_t nsamples, int intype, int outtype, const float* shifts, size_t nshifts,
                           uint32_t samplerate, uint32_t* samplenum, uint8_t* out, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    size_t ibps = intype == ORACLE_I16 ? 4 : 8, obps = outtype == ORACLE_I16 ? 4 : 8;
    size_t blk = ORACLE_BUFFER_SIZE / ibps;
    size_t nblocks = (nsamples + blk - 1) / blk;
    if (nblocks > nshifts) return -1.0;
    if ((size_t)threads > nblocks) threads = nblocks ? (int)nblocks : 1;
    oracle_blocks_slice* sl = (oracle_blocks_slice*)calloc((size_t)threads, sizeof(oracle_blocks_slice));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    size_t per = (nblocks + (size_t)threads - 1) / (size_t)threads;
    uint32_t sn = *samplenum;
    for (int t = 0; t < threads; t++) {
        size_t b0 = (size_t)t * per, b1 = b0 + per < nblocks ? b0 + per : nblocks;
        if (b0 > nblocks) b0 = nblocks;
        size_t k0 = b0 * blk, k1 = b1 * blk < nsamples ? b1 * blk : nsamples;
        if (k0 > nsamples) k0 = nsamples;
        sl[t] = (oracle_blocks_slice){in + k0 * ibps, out + k0 * obps, (k1 - k0) * ibps, intype, outtype, shifts + b0, nshifts - b0,
                                      samplerate, sn, 0};
        for (size_t b = b0; b < b1; b++) {   /* carry the state block by block, as the reference does */
            size_t n = (b + 1) * blk <= nsamples ? blk : nsamples - b * blk;
            sn = oracle_samplenum_advance(sn, shifts[b], samplerate, n);
        }
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, oracle_blocks_run, &sl[t]);
    oracle_blocks_run(&sl[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    long bad = 0;
    for (int t = 0; t < threads; t++) bad |= sl[t].rc < 0;
    free(sl);
    free(th);
    if (bad) return -1.0;
    *samplenum = sn;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------ */
/* decimating FIR fused after the mixer.  NOT in the reference (the usual  */
/* next block after `doppler` in an SDR pipe, README.md:53); this function IS its             */
/* specification, and the CUDA path must reproduce it bit for bit:                            */
/*   y[k]  = the mixer's Complex<f32> result for stream sample k (dsp.rs:117-134 exactly,     */
/*           per-block shifts as oracle_mix_blocks), y[k] = 0 for k < 0                       */
/*   z[m]  = sum over t = 0 .. ntaps-1, in that order, of h[t] * y[m*M - t], each step one     */
/*           fused multiply-add in f32 (fmaf), re and im separately, accumulator starting at 0 */
/*   out   = z as raw f32 pairs, or (z * 32767.0) as i16 like main.rs:73-87                    */
/* One output per M input samples: output m is produced by the call that supplies sample m*M. */
/* State across calls: the last ntaps-1 mixed samples (hist, oldest first), the stream        */
/* position *pos of the next input sample, and samplenum.  Returns bytes written or -1.      */
long oracle_mix_decimate(const uint8_t* in, size_t len, int intype, int outtype, const float* shifts, size_t nshifts,
                         uint32_t samplerate, uint32_t* samplenum, const float* taps, uint32_t ntaps, uint32_t M,
                         oracle_c32* hist, uint64_t* pos, uint8_t* out)
{
    size_t ibps = intype == ORACLE_I16 ? 4 : 8;
    if (len % ibps != 0 || ntaps == 0 || M == 0) return -1;
    size_t n = len / ibps, blk = ORACLE_BUFFER_SIZE / ibps, nh = ntaps - 1;
    if ((n + blk - 1) / blk > nshifts && n > 0) return -1;
    /* y with its history in front: line[0 .. nh) = hist, line[nh + i] = y of this call's sample i */
    oracle_c32* line = (oracle_c32*)malloc((nh + n + 1) * sizeof(oracle_c32));
    oracle_c32* a = (oracle_c32*)malloc((blk + 1) * sizeof(oracle_c32));
    memcpy(line, hist, nh * sizeof(oracle_c32));
    for (size_t k = 0, b = 0; k < n; k += blk, b++) {
        size_t c = n - k < blk ? n - k : blk;
        if (intype == ORACLE_I16)
            oracle_convert_iqi16_to_complex(in + k * ibps, c * ibps, a);
        else
            oracle_convert_iqf32_to_complex(in + k * ibps, c * ibps, a);
        oracle_shift_frequency(a, c, samplenum, shifts[b], samplerate, line + nh + k);
    }
    size_t wr = 0;
    uint64_t K0 = *pos;
    for (size_t i = 0; i < n; i++) {
        if ((K0 + i) % M != 0) continue;
        float re = 0.0f, im = 0.0f;
        for (uint32_t t = 0; t < ntaps; t++) {
            const oracle_c32 y = line[nh + i - t];   /* t <= nh: never before the history */
            re = fmaf(taps[t], y.re, re);
            im = fmaf(taps[t], y.im, im);
        }
        oracle_c32 z = {re, im};
        wr += outtype == ORACLE_I16 ? oracle_egress_i16(&z, 1, out + wr) : oracle_egress_f32(&z, 1, out + wr);
    }
    if (nh) memmove(hist, line + n, nh * sizeof(oracle_c32));   /* the last nh mixed samples (old history when n < nh) */
    *pos = K0 + n;
    free(line);
    free(a);
    return (long)wr;
}

/* Direct libm sincosf on a batch (checker for the product's device sincosf and its
 * host-compiled twin in tests/native). */
void oracle_sincosf_batch(const float* theta, size_t n, float* sin_out, float* cos_out)
{
    for (size_t i = 0; i < n; i++) sincosf(theta[i], &sin_out[i], &cos_out[i]);
}

/* theta_f32 = (-2*PI) * (shift_hz / fs as f32 * n as f32), dsp.rs:121. */
float oracle_theta(float shift_hz, uint32_t samplerate, uint32_t samplenum)
{
    const float PI_F32 = 3.14159274101257324219f;
    float x = shift_hz / (float)samplerate * (float)samplenum;
    float c = -2.0f * PI_F32;
    return c * x;
}

const char* oracle_libc_version(void)
{
    extern const char* gnu_get_libc_version(void);
    return gnu_get_libc_version();
}
