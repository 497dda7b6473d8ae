// mixer_kernels.cuh -- the sm_100a kernels of the doppler NCO-mixer hot path.
//
// One pass over HBM fuses what the reference does in three passes with three allocations per
// 8 KiB block (src/main.rs:65-94): ingest cast (src/dsp.rs:85-115), per-sample multiply by
// exp(-j*2*pi*(shift/fs)*samplenum) (src/dsp.rs:117-134), egress cast (src/main.rs:73-93).
//
// Layout in HBM: input and output are the reference's own interleaved little-endian IQ byte
// streams (4 B/sample i16, 8 B/sample f32), 16-byte aligned.  Work is cut into TILES of
// 256 threads x U groups x G samples; a thread's group is 8 or 16 contiguous bytes on the wide
// side, consecutive lanes take consecutive groups, so every warp-level load/store instruction
// covers one contiguous 256/512-byte span (fully coalesced, 128-bit where the format allows).
// Each CTA walks a contiguous run of tiles.
//
// The phase index is the reference's `samplenum` state machine, not the sample index.  The
// host planner (plan.h) supplies PIECES in which samplenum is closed-form; a tile that lies
// inside one piece takes a branch-free fast path, tiles that straddle pieces (or the ragged
// tail) take a per-sample path.  The phasor comes either from a TABLE of the piece's period
// (built once per shift by build_phasor_table_kernel with the same device routine, staged in
// shared memory when it fits, read through L1/L2 otherwise) or from direct evaluation of the
// bit-exact double-precision sincosf (sincosf_glibc.h).  A sin/cos recurrence is NOT used: the
// reference quantises theta to f32 before the trig call, which no recurrence reproduces.
//
// Roofline: HBM.  Algorithmic bytes per complex sample: i16->i16 8, i16->f32 12, f32->i16 12,
// f32->f32 16 (table / piece traffic is O(period) and excluded).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sincosf_glibc.h"

namespace dmix {

constexpr int kThreads = 256;   // default CTA size (table builder, converters, default mixer config)
constexpr int kUnroll = 4;      // default groups per thread per tile
constexpr uint32_t kNoTab = 0xffffffffu;
constexpr int kInlinePieces = 4;
constexpr int kTabPad = 4;   // table entries replicated past the period (>= max group size)

constexpr int I16 = 0;
constexpr int F32 = 1;

// samples per thread-group: the side with the wider sample gets 16 bytes per lane
__host__ __device__ constexpr int group_samples(int in, int out) { return (in == I16 && out == I16) ? 4 : 2; }
__host__ __device__ constexpr uint32_t tile_samples(int in, int out, int threads = kThreads, int unroll = kUnroll)
{
    return (uint32_t)threads * unroll * group_samples(in, out);
}

struct DevPiece {          // launch-relative sample indices
    uint32_t k_begin;
    uint32_t k_end;
    uint32_t base;         // linear: samplenum at k_begin; periodic: (samplenum - 1) at k_begin
    uint32_t period;       // 0 -> linear
    float r;               // shift_hz / f32(samplerate)
    uint32_t tab;          // first entry of this piece's phasor table in the arena, or kNoTab
    uint32_t magic;        // division by `period`: q = (x * magic) >> shift for x < 2^31
    uint32_t shift;
    uint32_t step_u;       // (CTA threads * G) mod period
    uint32_t pad[3];
};
static_assert(sizeof(DevPiece) == 48, "DevPiece layout");

struct MixArgs {
    const void* in;
    void* out;
    const DevPiece* pieces;   // global copy when npieces > kInlinePieces
    const float2* tables;     // phasor arena: entry = (cos, sin)
    uint32_t nsamples;
    uint32_t npieces;
    uint32_t ntiles;
    uint32_t tiles_per_cta;
    uint32_t smem_entries;    // capacity of the shared-memory table (excluding pad)
    uint32_t interleave;      // 0: CTA b owns tiles [b*tiles_per_cta, (b+1)*tiles_per_cta); 1: tiles b, b+grid, ...
    DevPiece inl[kInlinePieces];
};

// ---------------------------------------------------------------------------------------------
// phasor: ccexpf(0 + i*theta), theta = (-2*PI) * (r * f32(n))   (dsp.rs:121-122, complex.c:33-39)
// For a zero real part glibc's cexpf is exp(0)=1 times sincosf(theta); non-finite theta -> NaN.
__device__ __forceinline__ float2 phasor(float r, uint32_t n)
{
    const float x = __fmul_rn(r, __uint2float_rn(n));
    const float theta = __fmul_rn(__uint_as_float(0xC0C90FDBu) /* -2.0f * PI_f32 */, x);
    const db_sincos_t sc = db_sincosf_glibc(theta);
    return make_float2(sc.c, sc.s);
}

// sample * corrector, num-complex 0.1.35 Mul: (a*c - b*d, a*d + b*c), four products and two
// sums, each rounded separately (rustc never fuses).
__device__ __forceinline__ float2 cmul_unfused(float2 smp, float2 ph)
{
    const float ac = __fmul_rn(smp.x, ph.x);
    const float bd = __fmul_rn(smp.y, ph.y);
    const float ad = __fmul_rn(smp.x, ph.y);
    const float bc = __fmul_rn(smp.y, ph.x);
    return make_float2(__fsub_rn(ac, bd), __fadd_rn(ad, bc));
}

// dsp.rs:91-92: (i16 as f32) / 32768.   (exact: power-of-two scale)
__device__ __forceinline__ float2 ingest_i16(uint32_t w)
{
    const float i = __int2float_rn((int)(short)(w & 0xffffu));
    const float q = __int2float_rn((int)(short)(w >> 16));
    return make_float2(__fmul_rn(i, 0x1p-15f), __fmul_rn(q, 0x1p-15f));
}

// main.rs:77-78: (v * 32767.0) as i16 -- truncate toward zero, saturate, NaN -> 0: exactly
// PTX cvt.rzi.s16.f32.
__device__ __forceinline__ uint32_t egress_i16(float2 v)
{
    short i, q;
    const float fi = __fmul_rn(v.x, 32767.0f);
    const float fq = __fmul_rn(v.y, 32767.0f);
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(i) : "f"(fi));
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(q) : "f"(fq));
    return (uint32_t)(uint16_t)i | ((uint32_t)(uint16_t)q << 16);
}

// ---------------------------------------------------------------------------------------------
// streaming group loads / stores (evict-first: every byte is touched once)
template <int IN, int G>
__device__ __forceinline__ void load_group(const void* in, uint32_t g, float2 (&s)[G])
{
    if constexpr (IN == I16 && G == 4) {
        const uint4 w = __ldcs(reinterpret_cast<const uint4*>(in) + g);
        s[0] = ingest_i16(w.x);
        s[1] = ingest_i16(w.y);
        s[2] = ingest_i16(w.z);
        s[3] = ingest_i16(w.w);
    } else if constexpr (IN == I16 && G == 2) {
        const uint2 w = __ldcs(reinterpret_cast<const uint2*>(in) + g);
        s[0] = ingest_i16(w.x);
        s[1] = ingest_i16(w.y);
    } else {
        static_assert(G == 2, "f32 input groups are 2 samples");
        const float4 w = __ldcs(reinterpret_cast<const float4*>(in) + g);
        s[0] = make_float2(w.x, w.y);
        s[1] = make_float2(w.z, w.w);
    }
}

template <int OUT, int G>
__device__ __forceinline__ void store_group(void* out, uint32_t g, const float2 (&v)[G])
{
    if constexpr (OUT == I16 && G == 4) {
        __stcs(reinterpret_cast<uint4*>(out) + g,
               make_uint4(egress_i16(v[0]), egress_i16(v[1]), egress_i16(v[2]), egress_i16(v[3])));
    } else if constexpr (OUT == I16 && G == 2) {
        __stcs(reinterpret_cast<uint2*>(out) + g, make_uint2(egress_i16(v[0]), egress_i16(v[1])));
    } else {
        static_assert(G == 2, "f32 output groups are 2 samples");
        __stcs(reinterpret_cast<float4*>(out) + g, make_float4(v[0].x, v[0].y, v[1].x, v[1].y));
    }
}

template <int IN>
__device__ __forceinline__ float2 load_sample(const void* in, uint32_t k)
{
    if constexpr (IN == I16)
        return ingest_i16(__ldcs(reinterpret_cast<const uint32_t*>(in) + k));
    else
        return __ldcs(reinterpret_cast<const float2*>(in) + k);
}

template <int OUT>
__device__ __forceinline__ void store_sample(void* out, uint32_t k, float2 v)
{
    if constexpr (OUT == I16)
        __stcs(reinterpret_cast<uint32_t*>(out) + k, egress_i16(v));
    else
        __stcs(reinterpret_cast<float2*>(out) + k, v);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ DevPiece get_piece(const MixArgs& a, uint32_t i)
{
    if (a.npieces <= (uint32_t)kInlinePieces) return a.inl[i];   // kernel-parameter (constant) bank
    return a.pieces[i];
}

__device__ __forceinline__ uint32_t piece_end(const MixArgs& a, uint32_t i)
{
    if (a.npieces <= (uint32_t)kInlinePieces) return a.inl[i].k_end;
    return a.pieces[i].k_end;
}

// index of the piece containing sample k, searching forward from `from`
__device__ __forceinline__ uint32_t find_piece(const MixArgs& a, uint32_t from, uint32_t k)
{
    if (k < piece_end(a, from)) return from;
    if (from + 1 < a.npieces && k < piece_end(a, from + 1)) return from + 1;
    uint32_t lo = from + 1, hi = a.npieces - 1;   // invariant: answer in [lo, hi]
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (k < piece_end(a, mid))
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}

// samplenum of the sample at offset `off` from the piece start
__device__ __forceinline__ uint32_t piece_samplenum(const DevPiece& p, uint32_t off)
{
    if (p.period == 0) return p.base + off;
    const uint32_t x = p.base + off;
    const uint32_t q = (uint32_t)(((uint64_t)x * p.magic) >> p.shift);
    return x - q * p.period + 1u;
}

enum FastMode { kTabShared = 0, kTabGlobal = 1, kDirectPeriodic = 2, kDirectLinear = 3 };

// A full tile inside one piece.  All U group loads are issued before any arithmetic.
template <int IN, int OUT, int MODE, int T, int U>
__device__ __forceinline__ void fast_tile(const MixArgs& a, const DevPiece& p, uint32_t k0, const float2* tab)
{
    constexpr int G = group_samples(IN, OUT);
    const uint32_t g0 = k0 / G + threadIdx.x;
    float2 smp[U][G];
#pragma unroll
    for (int u = 0; u < U; u++) load_group<IN, G>(a.in, g0 + u * T, smp[u]);

    const uint32_t off = (k0 - p.k_begin) + threadIdx.x * G;
    uint32_t j = 0;
    if constexpr (MODE != kDirectLinear) j = piece_samplenum(p, off) - 1u;   // phase index in [0, period)

#pragma unroll
    for (int u = 0; u < U; u++) {
        float2 res[G];
#pragma unroll
        for (int s = 0; s < G; s++) {
            float2 ph;
            if constexpr (MODE == kTabShared) {
                ph = tab[j + s];                    // padded: no wrap inside a group
            } else if constexpr (MODE == kTabGlobal) {
                ph = __ldg(tab + j + s);
            } else if constexpr (MODE == kDirectPeriodic) {
                uint32_t n = j + s + 1u;
                if (n > p.period) n -= p.period;
                ph = phasor(p.r, n);
            } else {
                ph = phasor(p.r, p.base + off + (uint32_t)(u * T * G + s));
            }
            res[s] = cmul_unfused(smp[u][s], ph);
        }
        store_group<OUT, G>(a.out, g0 + u * T, res);
        if constexpr (MODE != kDirectLinear) {
            j += p.step_u;
            if (j >= p.period) j -= p.period;
        }
    }
}

// Generic per-sample tile: piece boundaries inside the tile and/or the ragged end of the buffer.
template <int IN, int OUT, int T, int U>
__device__ __noinline__ void slow_tile(const MixArgs& a, uint32_t pi, uint32_t k0)
{
    constexpr uint32_t kTile = tile_samples(IN, OUT, T, U);
    DevPiece p = get_piece(a, pi);
    for (uint32_t i = threadIdx.x; i < kTile; i += T) {
        const uint32_t k = k0 + i;
        if (k >= a.nsamples) break;
        if (k >= p.k_end) {
            pi = find_piece(a, pi, k);
            p = get_piece(a, pi);
        }
        const uint32_t n = piece_samplenum(p, k - p.k_begin);
        const float2 ph = phasor(p.r, n);
        store_sample<OUT>(a.out, k, cmul_unfused(load_sample<IN>(a.in, k), ph));
    }
}

// T threads per CTA, U groups per thread per tile, at least MINB resident CTAs per SM.
template <int IN, int OUT, int T = kThreads, int U = kUnroll, int MINB = 0>
__global__ void __launch_bounds__(T, MINB) mix_kernel(const __grid_constant__ MixArgs a)
{
    constexpr uint32_t kTile = tile_samples(IN, OUT, T, U);
    extern __shared__ float2 tab_s[];

    uint32_t tile, tile_end, tile_step;
    if (a.interleave) {
        tile = blockIdx.x;
        tile_end = a.ntiles;
        tile_step = gridDim.x;
    } else {
        tile = blockIdx.x * a.tiles_per_cta;
        tile_end = min(tile + a.tiles_per_cta, a.ntiles);
        tile_step = 1;
    }
    uint32_t pi = 0;
    uint32_t staged = 0xffffffffu;   // piece whose table is in shared memory
    DevPiece p = get_piece(a, 0);

    for (; tile < tile_end; tile += tile_step) {
        const uint32_t k0 = tile * kTile;
        if (k0 >= p.k_end) {
            pi = find_piece(a, pi, k0);
            p = get_piece(a, pi);
        }
        const bool fast = (k0 + kTile <= p.k_end) && (k0 + kTile <= a.nsamples);
        if (!fast) {
            slow_tile<IN, OUT, T, U>(a, pi, k0);
            continue;
        }
        if (p.period == 0) {
            fast_tile<IN, OUT, kDirectLinear, T, U>(a, p, k0, nullptr);
        } else if (p.tab == kNoTab) {
            fast_tile<IN, OUT, kDirectPeriodic, T, U>(a, p, k0, nullptr);
        } else if (p.period <= a.smem_entries) {
            if (staged != pi) {            // CTA-uniform
                __syncthreads();
                for (uint32_t e = threadIdx.x; e < p.period + kTabPad; e += T)
                    tab_s[e] = __ldg(a.tables + p.tab + e);
                __syncthreads();
                staged = pi;
            }
            fast_tile<IN, OUT, kTabShared, T, U>(a, p, k0, tab_s);
        } else {
            fast_tile<IN, OUT, kTabGlobal, T, U>(a, p, k0, a.tables + p.tab);
        }
    }
}

// Phasor table of one shift: entry j (0 <= j < period + kTabPad) = phasor(r, (j mod period) + 1).
__global__ void __launch_bounds__(kThreads) build_phasor_table_kernel(float2* tab, float r, uint32_t period, uint32_t entries)
{
    const uint32_t j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= entries) return;
    tab[j] = phasor(r, (j % period) + 1u);
}

// convert_iq{i16,f32}_to_complex alone (dsp.rs:85-115): ingest cast, no mixing.
template <int IN>
__global__ void __launch_bounds__(kThreads) convert_kernel(const void* in, float2* out, uint32_t nsamples)
{
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t k = blockIdx.x * kThreads + threadIdx.x; k < nsamples; k += stride)
        __stcs(out + k, load_sample<IN>(in, k));
}

// self-test probes (doppler_b200_phasor_probe / doppler_b200_sincosf_probe)
__global__ void phasor_probe_kernel(float r, uint32_t n0, uint32_t count, float* c, float* s)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float2 ph = phasor(r, n0 + i);
    c[i] = ph.x;
    s[i] = ph.y;
}

__global__ void sincosf_probe_kernel(uint32_t first, uint32_t stride, uint32_t count, float* s, float* c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const db_sincos_t r = db_sincosf_glibc(__uint_as_float(first + i * stride));
    s[i] = r.s;
    c[i] = r.c;
}

}  // namespace dmix
