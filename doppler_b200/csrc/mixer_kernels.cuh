// mixer_kernels.cuh -- the sm_100a kernels of the doppler NCO-mixer hot path.
//
// One pass over HBM fuses what the reference does in three passes with three allocations per
// 8 KiB block (src/main.rs:65-94): ingest cast (src/dsp.rs:85-115), per-sample multiply by
// exp(-j*2*pi*(shift/fs)*samplenum) (src/dsp.rs:117-134), egress cast (src/main.rs:73-93).
//
// Layout in HBM: input and output are the reference's own interleaved little-endian IQ byte
// streams (4 B/sample i16, 8 B/sample f32), 16-byte aligned.  Work is cut into TILES of U rows;
// a row is one GROUP per lane (G samples = 8 or 16 contiguous bytes on the wide side), consecutive
// lanes take consecutive groups.  Every warp is an independent cp.async.bulk pipeline over tiles
// (mix_grid_kernel: const mode; mix_stream_kernel: GRID and COLUMN segments, track mode).
//
// The phase index is the reference's `samplenum` state machine, not the sample index.  The
// host planner (plan.h) supplies PIECES in which samplenum is closed-form; a tile that lies
// inside one piece takes a branch-free fast path, tiles that straddle pieces take a per-sample
// path.  The phasor comes from a TABLE of the piece's period staged in shared memory (periods
// up to 4096, built once per shift by build_phasor_table_kernel with the same device routine),
// from a COLUMN window (long periods: evaluated once per column, reused over the rows of the
// period structure), or from direct evaluation per sample of the bit-exact double-precision
// sincosf (sincosf_glibc.h).  A sin/cos recurrence is NOT used: the reference quantises theta
// to f32 before the trig call, which no recurrence reproduces.
//
// Roofline: HBM.  Algorithmic bytes per complex sample: i16->i16 8, i16->f32 12, f32->i16 12,
// f32->f32 16 (table / piece traffic is O(period) and excluded).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "sincosf_glibc.h"

namespace dmix {

constexpr int kThreads = 256;   // CTA size of the table builder and the converters
constexpr uint32_t kNoTab = 0xffffffffu;
constexpr int kInlinePieces = 4;
constexpr int kTabPad = 4;   // table entries replicated past the period (>= max group size)

constexpr int I16 = 0;
constexpr int F32 = 1;

// samples per thread-group: the side with the wider sample gets 16 bytes per lane
__host__ __device__ constexpr int group_samples(int in, int out) { return (in == I16 && out == I16) ? 4 : 2; }

struct DevPiece {          // launch-relative sample indices
    uint32_t k_begin;
    uint32_t k_end;
    uint32_t base;         // linear: samplenum at k_begin; periodic: (samplenum - 1) at k_begin
    uint32_t period;       // 0 -> linear
    float r;               // shift_hz / f32(samplerate)
    uint32_t tab;          // first entry of this piece's phasor table in the arena, or kNoTab
    uint32_t magic;        // division by `period`: q = (x * magic) >> shift for x < 2^31
    uint32_t shift;
    uint32_t step_u;       // samples per warp-row (32 * G, a property of the type pair) mod period
    uint32_t pad[3];
};
static_assert(sizeof(DevPiece) == 48, "DevPiece layout");

// A launch is cut into SEGMENTS of two kinds (all sample indices launch-relative; segment bounds
// are multiples of the alignment granule, so every bulk copy is 16-byte aligned):
//   GRID    contiguous tiles of kTileSamples, one work unit per tile; may contain any pieces.
//   COLUMN  the granule-aligned inside [k_begin, k_end) of ONE periodic piece whose table does not
//           fit shared memory, seen as `rows` periods: row j holds phase 0 at k0 + j * period (k0
//           is signed: the first row may start before the piece) and owns the samples
//           [a_j, a_j+1) clipped to the segment, a_j = align_down(k0 + j * period).  A work unit is
//           one column tile c (phases c*T - s_j ... of every row) over `rows_per_unit` consecutive
//           rows: the warp evaluates the column's phasors ONCE (direct, bit-exact sincosf), parks
//           them in shared memory and reuses them for every row, so evaluation cost per sample
//           drops by the number of rows.  s_j = (k0 + j*period) - a_j (0 .. granule-1) is the
//           row's phase shift.
struct DevSeg {
    uint32_t unit_begin, unit_end;   // work units [unit_begin, unit_end)
    uint32_t k_begin, k_end;         // samples [k_begin, k_end)
    uint32_t piece;                  // GRID: piece containing k_begin; COLUMN: the periodic piece
    uint32_t rows;                   // COLUMN: periods touched (the first and last may be partial); 0 -> GRID
    uint32_t rows_per_unit;          // COLUMN: rows sharing one phasor evaluation
    uint32_t ncols;                  // COLUMN: column tiles per row
    uint32_t ncols_magic, ncols_shift;   // division by ncols (x < 2^31)
    uint32_t k0;                     // COLUMN: sample holding phase 0 of row 0 (int32: may be negative)
    uint32_t period;                 // COLUMN: copy of the piece's period
    float r;                         // COLUMN: copy of the piece's ratio
    uint32_t pad[3];
};
static_assert(sizeof(DevSeg) == 64, "DevSeg layout");
constexpr int kSegIndexShift = 6;   // MixArgs::seg_index has one entry per 64 work units

constexpr int kInlineSegs = 3;

// One tile of work, produced by lane 0's iterator when it issues the tile's bulk load and read
// back by the whole warp when the tile reaches the compute stage.
struct TileDesc {
    uint32_t k0;       // sample at tile offset 0 (int32: a clipped first tile may start before the buffer)
    uint32_t nsamp;    // tile offsets [skip, nsamp) are this tile's samples (granule multiples, nsamp <= kTileSamples);
                       // 0 = end of this warp's work
    uint32_t seg;      // segment index
    uint32_t info;     // COLUMN: kColFlag | kColFirst (evaluate the phasor window) | row shift s_j
    uint32_t phase0;   // COLUMN: phase of the column's first sample (c * kTileSamples); GRID: piece to search from
    uint32_t period;   // COLUMN: the piece's period
    float r;           // COLUMN: the piece's ratio
    uint32_t skip;     // leading tile offsets that belong to a neighbouring piece (COLUMN, first row only)
};
static_assert(sizeof(TileDesc) == 32, "TileDesc layout");
constexpr uint32_t kColFlag = 0x80000000u, kColFirst = 0x40000000u;
constexpr int kWinLead = 3;   // window entries ahead of the column's first phase (largest row shift)

struct MixArgs {
    const void* in;
    void* out;
    const DevPiece* pieces;   // global copy when npieces > kInlinePieces
    const DevSeg* segs;       // global copy when nsegs > kInlineSegs
    const uint32_t* seg_index;   // with segs: segment containing work unit (i << kSegIndexShift), one entry per 64 units
    uint32_t* unit_counter;      // segmented kernel: next unclaimed work unit (zeroed before the launch)
    const float2* tables;     // phasor arena: entry = (cos, sin)
    uint32_t nsamples;
    uint32_t npieces;
    uint32_t nsegs;
    uint32_t nunits;          // work units over all segments
    uint32_t tail_begin;      // samples [tail_begin, nsamples): the sub-granule end of the buffer
    uint32_t smem_piece;      // streaming kernel: piece whose table is staged in shared memory, or kNoPiece
    uint32_t plateau_scratch; // lean kernel: 1 when the launch carries a per-warp plateau scratch (direct-evaluation shape)
    uint32_t max_claim;       // segmented kernels: most work units claimed at once (guided self-scheduling), >= 1
    DevPiece inl[kInlinePieces];
    DevSeg inl_segs[kInlineSegs];
};

// ---------------------------------------------------------------------------------------------
// phasor: ccexpf(0 + i*theta), theta = (-2*PI) * (r * f32(n))   (dsp.rs:121-122, complex.c:33-39)
// For a zero real part glibc's cexpf is exp(0)=1 times sincosf(theta); non-finite theta -> NaN.
// Deliberately NOT inlined: this is the generic (any-range) routine of the rare paths -- tiles that
// straddle pieces / ranges / period wraps, table builds, probes; the hot paths use tables or the
// range-specialised rows below.  One shared copy keeps the kernels inside the instruction cache.
static __device__ __noinline__ float2 phasor(float r, uint32_t n)
{
    const float x = __fmul_rn(r, __uint2float_rn(n));
    const float theta = __fmul_rn(__uint_as_float(0xC0C90FDBu) /* -2.0f * PI_f32 */, x);
    const db_sincos_t sc = db_sincosf_glibc(theta);
    return make_float2(sc.c, sc.s);
}

// Packed fp32 (sm_100 FMUL2): two independent IEEE round-to-nearest multiplies per issued instruction.
// Only the MULTIPLIES are packed: ptxas contracts add.rn.f32x2 with a preceding mul.rn.f32x2 into FFMA2
// even under --fmad=false, which would break the reference's unfused arithmetic, so sums stay scalar.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// (1.0f, 1.0f), read from constant memory so that ptxas cannot see its value (below)
static __constant__ unsigned long long kOnesF32x2 = 0x3f8000003f800000ull;

// sample * corrector, num-complex 0.1.35 Mul: (a*c - b*d, a*d + b*c), four products and two
// sums, each rounded separately (rustc never fuses).  The products are formed pairwise,
// [a*c, b*c] and [b*(-d), a*d]; b*(-d) == -(b*d) and x + (-y) == x - y exactly in IEEE-754.
__device__ __forceinline__ float2 cmul_unfused(float2 smp, float2 ph)
{
    float ac, bc, nbd, ad;
    unpack_f32x2(mul_f32x2(pack_f32x2(smp.x, smp.y), pack_f32x2(ph.x, ph.x)), ac, bc);
    unpack_f32x2(mul_f32x2(pack_f32x2(smp.y, smp.x), pack_f32x2(-ph.y, ph.y)), nbd, ad);
    return make_float2(__fadd_rn(ac, nbd), __fadd_rn(ad, bc));
}

// The same product with the two sums as ONE packed instruction: fma(x, 1, y) rounds x * 1 + y == x + y once, exactly like the
// add.  It has to be an fma by a multiplier ptxas cannot see through (constant memory): add.rn.f32x2 -- and an fma by a literal
// 1.0 alike -- gets contracted with the preceding mul.rn.f32x2 into one FFMA2 even under --fmad=false, which drops the
// product's rounding.  Three issue slots per complex multiply instead of four.  Measured (profiles/r02_ab_cmul.md, same
// session, interleaved): it helps where instruction issue binds -- direct evaluation i16->i16 0.877 -> 0.901 of peak -- and
// costs 0.5-1.7 % where HBM binds (the packed FFMA2 shares the pipe of the FMUL2s, the scalar FADDs do not), so only the
// direct-evaluation rows and the fused decimator use it.
__device__ __forceinline__ float2 cmul_unfused_fma(float2 smp, float2 ph)
{
#ifdef DOPPLER_CMUL_TWO_FADD   // A/B build (make variant_cmul): two scalar adds everywhere
    return cmul_unfused(smp, ph);
#else
    const uint64_t acbc = mul_f32x2(pack_f32x2(smp.x, smp.y), pack_f32x2(ph.x, ph.x));
    const uint64_t nbdad = mul_f32x2(pack_f32x2(smp.y, smp.x), pack_f32x2(-ph.y, ph.y));
    float2 r;
    unpack_f32x2(fma_f32x2(nbdad, kOnesF32x2, acbc), r.x, r.y);
    return r;
#endif
}

// i16 -> f32 of the low / high half of an IQ word without the slow-pipe I2F.S16 (measured ~4.5 issue cycles per warp
// against ~1.7 for PRMT / SHF / I2FP, profiles/r01_pipes.jsonl): sign-extend in the integer pipe -- low half with one
// byte permute whose upper two selectors replicate the sign of byte 1 (PRMT 0x9910), high half with an arithmetic
// shift -- then I2FP.F32.S32.  Two instructions per component (the round-1 exponent-bias trick for the low half was three).
__device__ __forceinline__ float i16_lo_f32(uint32_t w)
{
    int v;
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(v) : "r"(w));   // {b0, b1, sign(b1), sign(b1)}: inline asm keeps ptxas from fusing into I2F.S16
    return __int2float_rn(v);
}
__device__ __forceinline__ float i16_hi_f32(uint32_t w) { return __int2float_rn((int)w >> 16); }

// dsp.rs:91-92: (i16 as f32) / 32768.   (exact: power-of-two scale)
__device__ __forceinline__ float2 ingest_i16(uint32_t w)
{
    return make_float2(__fmul_rn(i16_lo_f32(w), 0x1p-15f), __fmul_rn(i16_hi_f32(w), 0x1p-15f));
}

// main.rs:77-78: (v * 32767.0) as i16 -- truncate toward zero, saturate, NaN -> 0: exactly
// PTX cvt.rzi.s16.f32.
__device__ __forceinline__ uint32_t egress_i16(float2 v)
{
    short i, q;
    float fi, fq;
    unpack_f32x2(mul_f32x2(pack_f32x2(v.x, v.y), pack_f32x2(32767.0f, 32767.0f)), fi, fq);
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(i) : "f"(fi));
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(q) : "f"(fq));
    return (uint32_t)(uint16_t)i | ((uint32_t)(uint16_t)q << 16);
}

// ---------------------------------------------------------------------------------------------
// single-sample streaming loads / stores (ragged ends, converters)
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

enum FastMode { kTabShared = 0, kTabGlobal = 1 };   // phasor table staged in shared memory / read through L1-L2

// =============================================================================================
// Streaming kernel: every warp is an independent bulk-async (TMA 1-D) pipeline.
//
// Measured on B200 (profiles/r01_tune_*.jsonl): a register-staged LDG/STG loop tops out at
// 0.85-0.93 of the measured copy peak however it is shaped, while cp.async.bulk pipelines reach
// 0.99-1.02 when -- and only when -- about 32-48 KB of loads are in flight per SM (more in
// flight is *slower*: 0.93-0.95).  So the mixer moves its tiles with cp.async.bulk: global ->
// shared (mbarrier complete_tx), compute from shared into shared, shared -> global (bulk
// group).  Bytes in flight are set by WARPS x S x tile bytes, not by registers or occupancy.
//
// One CTA per SM, WARPS warps, no CTA-wide barrier after start-up.  Warp w of CTA b is
// pipeline p = b * WARPS + w and owns tiles p, p + npipes, ... (interleaved, so the whole chip
// walks one moving window of the stream).  Per tile and warp:
//     wait full[s]  ->  LDS the tile into registers  ->  lane 0: wait until the bulk store that
//     last used out[s] has drained  ->  syncwarp  ->  lane 0: refill in[s] with tile i + S  ->
//     multiply, STS into out[s]  ->  fence.proxy.async  ->  syncwarp  ->  lane 0: bulk store
// The shared-memory phasor table is de-interleaved into G planes (entry e -> plane e mod G) so
// that the warp's lookups (lane l needs entries j + G*l + s) are bank-conflict free.
constexpr uint32_t kNoPiece = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t cnt)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int IN, int OUT, int WARPS_, int S_, int U_>
struct StreamCfg {
    static constexpr int WARPS = WARPS_, S = S_, U = U_;
    static constexpr int G = group_samples(IN, OUT);
    static constexpr int kGroupIn = G * (IN == I16 ? 4 : 8);
    static constexpr int kGroupOut = G * (OUT == I16 ? 4 : 8);
    static constexpr int kRow = 32 * G;                     // samples per warp-row (one group per lane)
    static constexpr int kTileSamples = kRow * U;
    static constexpr int kTileIn = 32 * U * kGroupIn;
    static constexpr int kTileOut = 32 * U * kGroupOut;
    static constexpr int kRing = S * (kTileIn + kTileOut);  // per warp
    static constexpr int kBarBytes = ((WARPS * S * 8 + 127) / 128) * 128;
    static constexpr int kDescBytes = S * (int)sizeof(TileDesc);   // per warp
    // alignment granule in samples: both sides of every bulk copy must be 16-byte granular
    static constexpr int kGran = (IN == I16 || OUT == I16) ? 4 : 2;
    // COLUMN phasor window: kWinIters entries per lane cover phases c*T - kWinLead .. c*T + T;
    // G planes (entry e -> plane e mod G) of kWinPlane float2, kWinPlane = 16/G (mod 16) so that
    // the lane-consecutive 64-bit stores of the evaluation pass are bank-conflict free.
    static constexpr int kWinIters = (kTileSamples + kWinLead + 1 + 31) / 32;
    static constexpr int kWinPlane = ((32 * kWinIters / G + 15) / 16) * 16 + 16 / G;
    static constexpr int kWinBytes = G * kWinPlane * 8;            // per warp
    static constexpr int kFixedSmem = kBarBytes + WARPS * (kRing + kDescBytes + kWinBytes);   // mix_stream_kernel
    static constexpr int kGridSmem = kBarBytes + WARPS * kRing;                                // mix_grid_kernel (no descriptors / windows)
    // PLATEAU scratch (direct evaluation above samplenum 2^24, where f32(n) takes one value per 2 .. 256 consecutive n):
    // the distinct phasors of one tile, at most kTileSamples / 2 + 2 of them
    static constexpr int kPlateauEntries = ((kTileSamples / 2 + 2 + 15) / 16) * 16;
    static constexpr int kPlateauBytes = kPlateauEntries * 8;      // per warp; the segmented kernel reuses its COLUMN window instead
    static_assert(G * kWinPlane >= kPlateauEntries, "the COLUMN window doubles as the plateau scratch");
    // shared-memory table: G planes of plane_len(entries) float2 each
    __host__ __device__ static constexpr uint32_t plane_len(uint32_t period) { return (period + kRow + G - 1) / G + 1; }
    __host__ __device__ static constexpr uint32_t table_bytes(uint32_t period) { return G * plane_len(period) * 8; }
};

// i16 -> i16 only: the 2^-15 ingest scale (dsp.rs:91) is deferred into the egress constant.
// a = i * 2^-15 is exact, so fl(a*c) = fl(i*c) * 2^-15 and fl(re * 32767) = fl(re' * (32767 * 2^-15))
// bit for bit (power-of-two scaling commutes with rounding; the only exceptions are values
// below 2^-100, which truncate to 0 on the i16 egress either way).  Saves 2 FMUL per sample.
template <int IN, int OUT>
struct Scaling {
    static constexpr bool kDeferred = (IN == I16 && OUT == I16);
};

template <int IN, int OUT, int G>
__device__ __forceinline__ void unpack_group(const uint32_t (&w)[4], float2 (&s)[G])
{
    if constexpr (IN == I16) {
#pragma unroll
        for (int i = 0; i < G; i++) {
            if constexpr (Scaling<IN, OUT>::kDeferred) {
                s[i] = make_float2(i16_lo_f32(w[i]), i16_hi_f32(w[i]));
            } else {
                s[i] = ingest_i16(w[i]);
            }
        }
    } else {
        s[0] = make_float2(__uint_as_float(w[0]), __uint_as_float(w[1]));
        s[1] = make_float2(__uint_as_float(w[2]), __uint_as_float(w[3]));
    }
}

template <int IN, int OUT>
__device__ __forceinline__ uint32_t egress_i16_scaled(float2 v)
{
    if constexpr (Scaling<IN, OUT>::kDeferred) {
        short i, q;
        float fi, fq;
        unpack_f32x2(mul_f32x2(pack_f32x2(v.x, v.y), pack_f32x2(0x1.fffcp-1f, 0x1.fffcp-1f /* 32767 * 2^-15 */)), fi, fq);
        asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(i) : "f"(fi));
        asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(q) : "f"(fq));
        return (uint32_t)(uint16_t)i | ((uint32_t)(uint16_t)q << 16);
    } else {
        return egress_i16(v);
    }
}

// e in [0, period + kRow) -> e mod period (periods shorter than a row need the real modulo)
__device__ __forceinline__ uint32_t wrap_phase(uint32_t e, uint32_t period)
{
    if (e >= period) {
        e -= period;
        if (e >= period) e %= period;
    }
    return e;
}

// ---------------------------------------------------------------------------------------------
// Direct (table-free) evaluation, fast rows.
//
// The generic db_sincosf_glibc() costs ~115 issued instructions per sample inside the tile loop
// (three range branches, per-lane 4/pi window lookup, literal double constants re-materialised
// with two MOVs per use).  Inside one piece theta(n) = C * (r * f32(n)) is monotone in n (every
// step is a correctly-rounded multiply by a constant), so the first and last sample of a tile
// bound every sample between them: if both fall in the same glibc range -- and, in the large
// range, the same binade -- the whole tile does, warp-uniformly, and runs a branch-free
// evaluation of exactly that range's operation sequence:
//   LARGE  |theta| >= 120        : glibc reduce_large.  The 96-bit window of 4/pi depends only on
//          the exponent, so it is shifted once per tile (W = window << (e & 7), mod 2^96) and
//          m * W replaces (m << shift) * window -- the same integer mod 2^96, hence the same res0.
//   MEDIUM 0.75 <= |theta| < 120 : reduce_fast (one fused x - n*pi/2).
//   SMALL  2^-12 <= |theta| < .75: polynomial pair on theta itself.
//   TINY   |theta| < 2^-12       : (cos, sin) = (1, theta).
// The operation sequences (and their host twins) are in sincosf_glibc.h.
// Tiles that straddle ranges / binades / the period wrap take the generic per-sample routine.
enum DirectRange { kRangeGeneric = 0, kRangeLarge = 1, kRangeMedium = 2, kRangeSmall = 3, kRangeTiny = 4 };

// theta as the reference forms it (dsp.rs:121): f32(-2*PI) * (r * f32(n))
__device__ __forceinline__ float theta_of(float r, uint32_t n)
{
    return __fmul_rn(__uint_as_float(0xC0C90FDBu), __fmul_rn(r, __uint2float_rn(n)));
}

// Range of a tile whose samplenum runs n_first .. n_last without wrapping.
__device__ __forceinline__ int classify_tile(float r, uint32_t n_first, uint32_t n_last, db_window_t& t)
{
    const uint32_t b0 = __float_as_uint(theta_of(r, n_first)), b1 = __float_as_uint(theta_of(r, n_last));
    const uint32_t a0 = b0 & 0x7fffffffu, a1 = b1 & 0x7fffffffu;
    if (a0 > a1) return kRangeGeneric;   // cannot happen for finite r; NaN ordering is not relied on
    if (a1 < 0x39800000u) return kRangeTiny;
    if (a0 >= 0x39800000u && a1 < 0x3f400000u) return kRangeSmall;
    if (a0 >= 0x3f400000u && a1 < 0x42f00000u) return kRangeMedium;
    if (a0 >= 0x42f00000u && a1 < 0x7f800000u && (a0 >> 23) == (a1 >> 23)) {
        db_large_window(b0, &t);
        return kRangeLarge;
    }
    return kRangeGeneric;
}

template <int RANGE>
__device__ __forceinline__ float2 phasor_fast(float theta, const db_window_t& t)
{
    float s = theta, c = 1.0f;   // kRangeTiny: sin = theta, cos = 1
    if constexpr (RANGE == kRangeTiny) {
    } else if constexpr (RANGE == kRangeSmall) {
        db_sincosf_small(theta, &s, &c);
    } else if constexpr (RANGE == kRangeMedium) {
        db_sincosf_medium(theta, &s, &c);
    } else {
        db_sincosf_large(__float_as_uint(theta), &t, &s, &c);
    }
    return make_float2(c, s);
}

// Writes one lane's group of row u into the warp's output stage.
template <typename C, int IN, int OUT>
__device__ __forceinline__ void stage_group(unsigned char* out_s, int u, uint32_t lane, const float2 (&res)[C::G])
{
    unsigned char* dst = out_s + (u * 32 + lane) * C::kGroupOut;
    if constexpr (OUT == I16 && C::G == 4) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(egress_i16_scaled<IN, OUT>(res[0]), egress_i16_scaled<IN, OUT>(res[1]),
                                                    egress_i16_scaled<IN, OUT>(res[2]), egress_i16_scaled<IN, OUT>(res[3]));
    } else if constexpr (OUT == I16) {
        *reinterpret_cast<uint2*>(dst) = make_uint2(egress_i16(res[0]), egress_i16(res[1]));
    } else {
        *reinterpret_cast<float4*>(dst) = make_float4(res[0].x, res[0].y, res[1].x, res[1].y);
    }
}

// One full tile inside one piece, from the warp's shared-memory stage; phasors from a table.
template <typename C, int IN, int OUT, int MODE>
__device__ __forceinline__ void stream_tile(const uint32_t (&raw)[C::U][4], unsigned char* out_s, const DevPiece& p, uint32_t k0,
                                            const float2* tab, uint32_t plane_len, uint32_t lane)
{
    static_assert(MODE == kTabShared || MODE == kTabGlobal, "table modes only");
    constexpr int G = C::G, U = C::U;
    // phase index of the row's first sample (warp-uniform); lanes add G * lane
    uint32_t j = piece_samplenum(p, k0 - p.k_begin) - 1u;
#pragma unroll
    for (int u = 0; u < U; u++) {
        float2 smp[G], res[G];
        unpack_group<IN, OUT, G>(raw[u], smp);
#pragma unroll
        for (int s = 0; s < G; s++) {
            float2 ph;
            if constexpr (MODE == kTabShared) {
                const uint32_t e = j + (uint32_t)s;                          // uniform part of the entry index
                ph = tab[(e % G) * plane_len + e / G + lane];                // entry e + G*lane, plane (e mod G)
            } else {
                ph = __ldg(tab + wrap_phase(j + lane * G + (uint32_t)s, p.period));
            }
            res[s] = cmul_unfused(smp[s], ph);
        }
        stage_group<C, IN, OUT>(out_s, u, lane, res);
        j += p.step_u;
        if (j >= p.period) j -= p.period;
    }
}

// All rows of a tile with the phasor of tile-relative sample i given by ph(i).
template <typename C, int IN, int OUT, typename PH>
__device__ __forceinline__ void stream_rows(const uint32_t (&raw)[C::U][4], unsigned char* out_s, uint32_t lane, PH ph)
{
    constexpr int G = C::G, U = C::U;
#pragma unroll
    for (int u = 0; u < U; u++) {
        float2 smp[G], res[G];
        unpack_group<IN, OUT, G>(raw[u], smp);
#pragma unroll
        for (int s = 0; s < G; s++) res[s] = cmul_unfused_fma(smp[s], ph(u * C::kRow + s));
        stage_group<C, IN, OUT>(out_s, u, lane, res);
    }
}

// One full tile inside one table-less piece (linear, or periodic without a table): direct evaluation.
//
// PLATEAUS.  The reference forms theta from `samplenum as f32` (dsp.rs:121), and above 2^24 that conversion takes ONE
// value for 2, 4, ... 256 consecutive samplenums (round to nearest even): a stream that never resets (tiny |r|) spends
// almost all of its life there.  Inside one tile the distinct values are the floats between f32(n_first) and
// f32(n_last) -- consecutive bit patterns, so value e is bits(f32(n_first)) + e and sample n uses entry
// bits(f32(n)) - bits(f32(n_first)).  The warp evaluates each distinct phasor once into `scratch` (<= T/2 + 1
// evaluations for T samples, T/256 + 1 near the top of the u32 range) and every sample looks its own up.
template <typename C, int IN, int OUT>
__device__ __forceinline__ void stream_tile_direct(const uint32_t (&raw)[C::U][4], unsigned char* out_s, const DevPiece& p,
                                                   uint32_t k0, uint32_t lane, float2* scratch)
{
    constexpr uint32_t kTile = C::kTileSamples;
    const uint32_t off0 = k0 - p.k_begin;
    const float r = p.r;
    uint32_t n0;   // samplenum of the tile's first sample
    bool mono;     // samplenum runs n0, n0 + 1, ... through the whole tile (no period / u32 wrap)
    if (p.period == 0) {
        n0 = p.base + off0;
        mono = n0 + (kTile - 1u) >= n0;
    } else {
        const uint32_t j = piece_samplenum(p, off0) - 1u;
        n0 = j + 1u;
        mono = j + kTile <= p.period;
    }
    db_window_t dt;
    const int range = mono ? classify_tile(r, n0, n0 + (kTile - 1u), dt) : (int)kRangeGeneric;
    const uint32_t nl = n0 + lane * C::G;   // this lane's first samplenum in row 0
    // Only for i16 output: those pairs move 8-12 bytes per sample and direct evaluation makes them issue-bound, so fewer
    // evaluations win (0.755 -> 0.89 of the HBM peak for i16->i16).  With f32 output the per-sample evaluation already fits
    // under the HBM time, and the evaluate -> park -> barrier -> look up chain only adds exposed latency (measured 0.97 -> 0.90).
    if (OUT == I16 && mono && scratch != nullptr && n0 >= (1u << 24)) {
        const uint32_t b0 = __float_as_uint(__uint2float_rn(n0));
        const uint32_t nvals = __float_as_uint(__uint2float_rn(n0 + (kTile - 1u))) - b0 + 1u;   // <= kTile / 2 + 1
        auto theta_e = [&](uint32_t e) { return __fmul_rn(__uint_as_float(0xC0C90FDBu), __fmul_rn(r, __uint_as_float(b0 + e))); };
        if (range == kRangeLarge) {
#pragma unroll 2
            for (uint32_t e = lane; e < nvals; e += 32) scratch[e] = phasor_fast<kRangeLarge>(theta_e(e), dt);
        } else if (range == kRangeMedium) {
#pragma unroll 2
            for (uint32_t e = lane; e < nvals; e += 32) scratch[e] = phasor_fast<kRangeMedium>(theta_e(e), dt);
        } else if (range == kRangeSmall) {
#pragma unroll 2
            for (uint32_t e = lane; e < nvals; e += 32) scratch[e] = phasor_fast<kRangeSmall>(theta_e(e), dt);
        } else if (range == kRangeTiny) {
            for (uint32_t e = lane; e < nvals; e += 32) scratch[e] = phasor_fast<kRangeTiny>(theta_e(e), dt);
        } else {
#pragma unroll 1
            for (uint32_t e = lane; e < nvals; e += 32) {
                const db_sincos_t sc = db_sincosf_glibc(theta_e(e));
                scratch[e] = make_float2(sc.c, sc.s);
            }
        }
        __syncwarp();
        // (the caller's __syncwarp before the tile's bulk store separates these reads from the next tile's writes)
        stream_rows<C, IN, OUT>(raw, out_s, lane, [&](int i) { return scratch[__float_as_uint(__uint2float_rn(nl + (uint32_t)i)) - b0]; });
        return;
    }
    if (range == kRangeLarge) {
        stream_rows<C, IN, OUT>(raw, out_s, lane, [&](int i) { return phasor_fast<kRangeLarge>(theta_of(r, nl + (uint32_t)i), dt); });
    } else if (range == kRangeMedium) {
        stream_rows<C, IN, OUT>(raw, out_s, lane, [&](int i) { return phasor_fast<kRangeMedium>(theta_of(r, nl + (uint32_t)i), dt); });
    } else if (range == kRangeSmall) {
        stream_rows<C, IN, OUT>(raw, out_s, lane, [&](int i) { return phasor_fast<kRangeSmall>(theta_of(r, nl + (uint32_t)i), dt); });
    } else if (range == kRangeTiny) {
        stream_rows<C, IN, OUT>(raw, out_s, lane, [&](int i) { return phasor_fast<kRangeTiny>(theta_of(r, nl + (uint32_t)i), dt); });
    } else if (p.period == 0) {
        stream_rows<C, IN, OUT>(raw, out_s, lane, [&](int i) { return phasor(r, nl + (uint32_t)i); });
    } else {
        const uint32_t period = p.period;
        stream_rows<C, IN, OUT>(raw, out_s, lane, [&](int i) {
            uint32_t e = nl - 1u + (uint32_t)i;
            if (e >= period) e %= period;
            return phasor(r, e + 1u);
        });
    }
}

// A tile that straddles pieces: per-sample piece lookup and generic phasor, still staged.
template <typename C, int IN, int OUT>
__device__ __noinline__ void stream_tile_slow(const MixArgs& a, uint32_t pi, const unsigned char* in_s, unsigned char* out_s,
                                              uint32_t k0, uint32_t nsamp, uint32_t lane)
{
    DevPiece p = get_piece(a, pi);
    for (uint32_t i = lane; i < nsamp; i += 32) {
        const uint32_t k = k0 + i;
        if (k >= p.k_end) {
            pi = find_piece(a, pi, k);
            p = get_piece(a, pi);
        }
        const float2 ph = phasor(p.r, piece_samplenum(p, k - p.k_begin));
        float2 smp;
        if constexpr (IN == I16)
            smp = ingest_i16(reinterpret_cast<const uint32_t*>(in_s)[i]);
        else
            smp = reinterpret_cast<const float2*>(in_s)[i];
        const float2 v = cmul_unfused(smp, ph);
        if constexpr (OUT == I16)
            reinterpret_cast<uint32_t*>(out_s)[i] = egress_i16(v);
        else
            reinterpret_cast<float2*>(out_s)[i] = v;
    }
}

__host__ __device__ __forceinline__ DevSeg get_seg(const MixArgs& a, uint32_t i)
{
    if (a.nsegs <= (uint32_t)kInlineSegs) return a.inl_segs[i];
    return a.segs[i];
}

// COLUMN: evaluate the phasor window of one column into the warp's planes.  Entry e holds the
// phasor of phase (phase0 - kWinLead + e) mod period, i.e. samplenum = that + 1.
template <typename C>
__device__ __forceinline__ void eval_window(float2* win, float r, uint32_t period, uint32_t phase0, uint32_t lane)
{
    constexpr int G = C::G, NW = C::kWinIters, PL = C::kWinPlane;
    // phase of entry 0; only column 0 reaches back into the previous period
    const uint32_t f0 = phase0 >= (uint32_t)kWinLead ? phase0 - kWinLead : phase0 + period - kWinLead;
    const bool mono = f0 + 32u * NW <= period;   // no period wrap inside the window
    db_window_t dt;
    const int range = mono ? classify_tile(r, f0 + 1u, f0 + 32u * NW, dt) : (int)kRangeGeneric;
    float2* dst = win + (lane % G) * PL + lane / G;   // entry e = lane + 32*it -> plane e % G, slot e / G
    const uint32_t nl = f0 + 1u + lane;
    if (range == kRangeLarge) {
#pragma unroll 3
        for (int it = 0; it < NW; it++) dst[it * (32 / G)] = phasor_fast<kRangeLarge>(theta_of(r, nl + 32u * it), dt);
    } else if (range == kRangeMedium) {
#pragma unroll 3
        for (int it = 0; it < NW; it++) dst[it * (32 / G)] = phasor_fast<kRangeMedium>(theta_of(r, nl + 32u * it), dt);
    } else if (range == kRangeSmall) {
#pragma unroll 3
        for (int it = 0; it < NW; it++) dst[it * (32 / G)] = phasor_fast<kRangeSmall>(theta_of(r, nl + 32u * it), dt);
    } else if (range == kRangeTiny) {
#pragma unroll 3
        for (int it = 0; it < NW; it++) dst[it * (32 / G)] = phasor_fast<kRangeTiny>(theta_of(r, nl + 32u * it), dt);
    } else {
#pragma unroll 1
        for (int it = 0; it < NW; it++) {
            uint32_t f = f0 + lane + 32u * it;
            if (f >= period) f -= period;
            if (f >= period) f %= period;
            dst[it * (32 / G)] = phasor(r, f + 1u);
        }
    }
}

// COLUMN: one row tile against the parked window; sh = the row's phase shift s_j.
template <typename C, int IN, int OUT, int PL = C::kWinPlane>
__device__ __forceinline__ void stream_tile_column(const uint32_t (&raw)[C::U][4], unsigned char* out_s, const float2* win,
                                                   uint32_t sh, uint32_t lane)
{
    constexpr int G = C::G, U = C::U;
    constexpr int LOGG = G == 4 ? 2 : 1;
#pragma unroll
    for (int u = 0; u < U; u++) {
        float2 smp[G], res[G];
        unpack_group<IN, OUT, G>(raw[u], smp);
#pragma unroll
        for (int g = 0; g < G; g++) {
            // sample u*32G + G*lane + g has phase phase0 + that - sh -> window entry that + kWinLead - sh
            const uint32_t x = (uint32_t)(g + kWinLead) - sh;                     // warp-uniform, 0 .. G + 2
            const float2 ph = win[(x & (G - 1)) * PL + (x >> LOGG) + u * 32 + lane];
            res[g] = cmul_unfused(smp[g], ph);
        }
        stage_group<C, IN, OUT>(out_s, u, lane, res);
    }
}

// Work units -> tiles.  Work units are claimed one at a time from a global counter (a COLUMN unit costs 2-64 tiles plus a
// window evaluation, a GRID unit kGridUnitTiles tiles: a static split left ~10 % of the SMs idle at the end of a track-mode
// launch).  Every tile of a unit is a CLOSED FORM of (segment, unit, tile index) -- unit_tile() below -- so nothing about the
// expansion is sequential: on the device all 32 lanes of the warp expand the unit's tiles at once, lane t tile t (TileSrc),
// and handing out the next tile is a ballot and a few shuffles instead of a scalar iterator run by lane 0 (round 1: ~55
// issued instructions per tile with one active lane).  On the host (doppler_b200_plan_tiles_trace) TileIter walks the same
// closed form tile by tile, pipeline p taking units p, p + npipes, ...
constexpr uint32_t kGridUnitTiles = 8;   // a GRID work unit is this many consecutive tiles (comparable to a COLUMN unit)

struct UnitPos {        // warp-uniform position of one work unit inside its segment
    uint32_t ntiles;    // GRID: tiles of the unit; COLUMN: rows of the unit (rows whose column tile is empty are skipped)
    uint32_t c;         // COLUMN: column tile
    uint32_t j0;        // GRID: first tile of the unit within the segment; COLUMN: first row
};

template <typename C>
__host__ __device__ __forceinline__ UnitPos unit_pos(const DevSeg& sg, uint32_t u)
{
    constexpr uint32_t T = C::kTileSamples;
    const uint32_t v = u - sg.unit_begin;
    UnitPos up;
    if (sg.rows == 0) {
        const uint32_t ntiles = (sg.k_end - sg.k_begin + T - 1) / T;
        up.c = 0;
        up.j0 = v * kGridUnitTiles;
        up.ntiles = up.j0 + kGridUnitTiles < ntiles ? kGridUnitTiles : ntiles - up.j0;
    } else {
        const uint32_t g = (uint32_t)(((uint64_t)v * sg.ncols_magic) >> sg.ncols_shift);
        up.c = v - g * sg.ncols;
        up.j0 = g * sg.rows_per_unit;
        up.ntiles = up.j0 + sg.rows_per_unit < sg.rows ? sg.rows_per_unit : sg.rows - up.j0;
    }
    return up;
}

// Tile t of the unit: k0 / nsamp / skip / info (row shift, kColFlag); nsamp == 0 when the row has no samples in this column.
template <typename C>
__host__ __device__ __forceinline__ void unit_tile(const DevSeg& sg, const UnitPos& up, uint32_t t, uint32_t& k0, uint32_t& nsamp,
                                                   uint32_t& skip, uint32_t& info)
{
    constexpr uint32_t T = C::kTileSamples, GM = ~(uint32_t)(C::kGran - 1);
    if (sg.rows == 0) {
        const uint32_t start = sg.k_begin + (up.j0 + t) * T;
        k0 = start;
        nsamp = sg.k_end - start < T ? sg.k_end - start : T;
        skip = 0;
        info = 0;
        return;
    }
    const uint32_t j = up.j0 + t;
    const int32_t kj = (int32_t)sg.k0 + (int32_t)(j * sg.period);
    const int32_t aj = kj & (int32_t)GM, aj1 = (kj + (int32_t)sg.period) & (int32_t)GM;
    const int32_t start = aj + (int32_t)(up.c * T);
    int32_t hi = start + (int32_t)T < aj1 ? start + (int32_t)T : aj1;   // the row's last column is short
    if (hi > (int32_t)sg.k_end) hi = (int32_t)sg.k_end;                   // the segment's last row is partial
    const int32_t lo = start > (int32_t)sg.k_begin ? start : (int32_t)sg.k_begin;   // so is its first
    k0 = (uint32_t)start;
    skip = lo < hi ? (uint32_t)(lo - start) : 0u;
    nsamp = lo < hi ? (uint32_t)(hi - start) : 0u;
    info = kColFlag | (uint32_t)(kj - aj);
}

// Host walk (tests): pipeline `pipe` of `npipes` takes units pipe, pipe + npipes, ... and expands each tile by tile.
template <typename C>
struct TileIter {
    uint32_t u, step, nunits, seg, t;
    DevSeg sg;
    UnitPos up;
    bool first;

    __host__ void init(const MixArgs& a, uint32_t pipe, uint32_t npipes)
    {
        step = npipes;
        nunits = a.nunits;
        u = pipe;
        seg = 0;
        sg = get_seg(a, 0);
        up.ntiles = up.c = up.j0 = 0;
        t = 0;
        first = false;
    }

    __host__ bool next(const MixArgs& a, TileDesc& d)
    {
        for (;;) {
            if (t < up.ntiles) {
                unit_tile<C>(sg, up, t++, d.k0, d.nsamp, d.skip, d.info);
                if (d.nsamp == 0) continue;
                d.seg = seg;
                d.phase0 = sg.rows ? up.c * (uint32_t)C::kTileSamples : sg.piece;
                d.period = sg.rows ? sg.period : 0u;
                d.r = sg.rows ? sg.r : 0.0f;
                if (sg.rows && first) d.info |= kColFirst;
                first = false;
                return true;
            }
            if (u >= nunits) return false;
            while (u >= sg.unit_end) sg = get_seg(a, ++seg);
            up = unit_pos<C>(sg, u);
            u += step;
            t = 0;
            first = true;
        }
    }
};

#if defined(__CUDACC__)
// Device tile source: every member is executed by ALL lanes of the warp (warp-convergent); results are warp-uniform.
// Units are claimed one at a time (MixArgs::max_claim = 1), the claim for the NEXT unit issued when a unit starts so that its
// latency hides behind the unit's tiles.  Larger, guided claims (up to max_claim consecutive units while the launch is young,
// single units near the end) save atomics and segment lookups but measured 6-15 % SLOWER (profiles/r02_ab_seg.jsonl); the
// mechanism stays for such A/B runs.
constexpr uint32_t kMaxClaim = 8;
template <typename C>
struct TileSrc {
    uint32_t nunits, npipes, max_claim;
    uint32_t pending, pending_cnt;     // pending (lane 0): first unit of the chunk claimed ahead; pending_cnt: its size
    uint32_t u_cur, u_end;             // units left in the current chunk
    uint32_t seg, seg_unit_end;        // current segment
    UnitPos up;
    uint32_t t_next, batch;            // next tile of the unit to hand out; tile index held by lane 0
    uint32_t phase0, period;
    float r;
    bool first;
    uint32_t dk0, dns, dskip, dinfo;   // lane-held: tile (batch + lane) of the unit

    __device__ __forceinline__ uint32_t claim_size(uint32_t claimed) const
    {
        const uint32_t left = claimed < nunits ? nunits - claimed : 0u;
        const uint32_t c = left / (4u * npipes);
        return c < 1u ? 1u : (c > max_claim ? max_claim : c);
    }

    __device__ __forceinline__ void init(const MixArgs& a, uint32_t lane, uint32_t npipes_)
    {
        nunits = a.nunits;
        npipes = npipes_;
        max_claim = a.max_claim ? a.max_claim : 1u;
        pending_cnt = claim_size(0);
        pending = lane == 0 ? atomicAdd(a.unit_counter, pending_cnt) : 0u;
        u_cur = u_end = 0;
        seg = 0;
        seg_unit_end = get_seg(a, 0).unit_end;
        up.ntiles = up.c = up.j0 = 0;
        t_next = batch = 0;
        phase0 = period = 0;
        r = 0.0f;
        first = false;
        dk0 = dns = dskip = dinfo = 0;
    }

    __device__ __forceinline__ void expand(const DevSeg& sg, uint32_t lane)
    {
        const uint32_t t = batch + lane;
        dns = 0;
        if (t < up.ntiles) unit_tile<C>(sg, up, t, dk0, dns, dskip, dinfo);
    }

    // Next tile of this warp's stream of work; false when the launch has no more work for it.
    __device__ __forceinline__ bool fetch(const MixArgs& a, uint32_t lane, TileDesc& d)
    {
        for (;;) {
            if (t_next < up.ntiles) {
                if (t_next >= batch + 32u) {   // units of more than 32 rows: the next 32
                    batch += 32u;
                    expand(get_seg(a, seg), lane);
                }
                const uint32_t m = __ballot_sync(0xffffffffu, dns != 0u) & (0xffffffffu << (t_next - batch));
                if (m == 0u) {
                    t_next = batch + 32u;
                    continue;
                }
                const int l = __ffs((int)m) - 1;
                t_next = batch + (uint32_t)l + 1u;
                d.k0 = __shfl_sync(0xffffffffu, dk0, l);
                d.nsamp = __shfl_sync(0xffffffffu, dns, l);
                d.skip = __shfl_sync(0xffffffffu, dskip, l);
                d.info = __shfl_sync(0xffffffffu, dinfo, l) | (first && period ? kColFirst : 0u);
                d.seg = seg;
                d.phase0 = phase0;
                d.period = period;
                d.r = r;
                first = false;
                return true;
            }
            if (u_cur >= u_end) {   // chunk exhausted: take the one claimed earlier, claim the one after
                const uint32_t base = __shfl_sync(0xffffffffu, pending, 0);
                if (base >= nunits) return false;
                u_cur = base;
                u_end = base + pending_cnt < nunits ? base + pending_cnt : nunits;
                pending_cnt = claim_size(u_end);
                if (lane == 0) pending = atomicAdd(a.unit_counter, pending_cnt);
            }
            const uint32_t u = u_cur++;
            if (u >= seg_unit_end) {
                // jump through the coarse unit -> segment index, then walk the last few segments
                if (a.nsegs > (uint32_t)kInlineSegs) {
                    const uint32_t hint = a.seg_index[u >> kSegIndexShift];
                    if (hint > seg) seg = hint - 1;
                }
                do seg_unit_end = get_seg(a, ++seg).unit_end;
                while (u >= seg_unit_end);
            }
            const DevSeg sg = get_seg(a, seg);
            up = unit_pos<C>(sg, u);
            phase0 = sg.rows ? up.c * (uint32_t)C::kTileSamples : sg.piece;
            period = sg.rows ? sg.period : 0u;
            r = sg.rows ? sg.r : 0.0f;
            batch = 0;
            t_next = 0;
            first = true;
            expand(sg, lane);
        }
    }
};
#endif

template <int IN, int OUT, int WARPS, int S, int U>
__global__ void __launch_bounds__(WARPS * 32, 1) mix_stream_kernel(const __grid_constant__ MixArgs a)
{
    using C = StreamCfg<IN, OUT, WARPS, S, U>;
    constexpr int G = C::G;
    constexpr uint32_t kInBps = IN == I16 ? 4 : 8, kOutBps = OUT == I16 ? 4 : 8;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* rings = smem + C::kBarBytes;
    unsigned char* descs = rings + WARPS * C::kRing;
    unsigned char* wins = descs + WARPS * C::kDescBytes;
    float2* tab_s = reinterpret_cast<float2*>(smem + C::kFixedSmem);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* full = bars + warp * S;
    unsigned char* ring_in = rings + warp * C::kRing;
    unsigned char* ring_out = ring_in + S * C::kTileIn;
    TileDesc* desc_ring = reinterpret_cast<TileDesc*>(descs + warp * C::kDescBytes);
    float2* win = reinterpret_cast<float2*>(wins + warp * C::kWinBytes);

    // start-up: every warp owns its barriers (lane 0 initialises them, issues on them, the warp waits on them),
    // so the first loads go out before the CTA stages the table: their latency hides the staging
    if (lane == 0) {
        for (int s = 0; s < S; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const uint32_t pipe = blockIdx.x * WARPS + warp, npipes = gridDim.x * WARPS;
    const unsigned char* gin = static_cast<const unsigned char*>(a.in);
    unsigned char* gout = static_cast<unsigned char*>(a.out);

    // the warp expands its work units into tiles together (TileSrc); lane 0 issues the tile's bulk load and parks the
    // descriptor in shared memory for the compute stage
    TileSrc<C> src;
    src.init(a, lane, npipes);
    auto issue_next = [&](uint32_t s) {   // all lanes (warp-convergent)
        TileDesc d;
        const bool more = src.fetch(a, lane, d);
        if (lane == 0) {
            if (more) {
                const uint32_t bytes = (d.nsamp - d.skip) * kInBps;
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(ring_in + s * C::kTileIn + d.skip * kInBps, gin + (size_t)(d.k0 + d.skip) * kInBps, bytes, &full[s]);
            } else {
                d.k0 = d.nsamp = d.seg = d.info = d.phase0 = d.period = d.skip = 0;
                d.r = 0.0f;
            }
            desc_ring[s] = d;
        }
    };
    for (uint32_t s = 0; s < (uint32_t)S; s++) issue_next(s);

    // (optionally) one piece's table de-interleaved into shared memory
    uint32_t plane_len = 0;
    if (a.smem_piece != kNoPiece) {
        const DevPiece sp = get_piece(a, a.smem_piece);
        plane_len = C::plane_len(sp.period);
        const float2* src = a.tables + sp.tab;
        for (uint32_t e = threadIdx.x; e < sp.period + (uint32_t)C::kRow; e += WARPS * 32)
            tab_s[(e % G) * plane_len + e / G] = __ldg(src + e % sp.period);   // replicated past the period: no wrap in a row
    }
    __syncthreads();

    uint32_t pi = 0;                 // GRID: cached piece
    DevPiece p = get_piece(a, 0);
    for (uint32_t i = 0;; i++) {
        const uint32_t s = i % S;
        __syncwarp();
        const TileDesc d = desc_ring[s];
        if (d.nsamp == 0) break;
        const unsigned char* in_s = ring_in + s * C::kTileIn;
        unsigned char* out_s = ring_out + s * C::kTileOut;
        mbar_wait(&full[s], (i / S) & 1u);
        uint32_t raw[U][4];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned char* src = in_s + (u * 32 + lane) * C::kGroupIn;
            if constexpr (C::kGroupIn == 16) {
                const uint4 w = *reinterpret_cast<const uint4*>(src);
                raw[u][0] = w.x, raw[u][1] = w.y, raw[u][2] = w.z, raw[u][3] = w.w;
            } else {
                const uint2 w = *reinterpret_cast<const uint2*>(src);
                raw[u][0] = w.x, raw[u][1] = w.y, raw[u][2] = 0, raw[u][3] = 0;
            }
        }
        if (lane == 0) bulk_wait_read<S - 1>();   // the store that last read out[s] (tile i - S) has drained
        __syncwarp();                              // in[s] and desc[s] are in registers now
        bool refilled = false;
        if (d.info & kColFlag) {
            issue_next(s);
            refilled = true;
            if (d.info & kColFirst) {
                eval_window<C>(win, d.r, d.period, d.phase0, lane);
                __syncwarp();
            }
            stream_tile_column<C, IN, OUT>(raw, out_s, win, d.info & 0xffu, lane);
        } else {
            const uint32_t k0 = d.k0;
            if (k0 >= p.k_end || k0 < p.k_begin) {
                // the segment names the piece holding its first sample: walk forward from there (GRID
                // segments between COLUMN segments hold a handful of pieces)
                pi = max(d.phase0, k0 >= p.k_end ? pi : 0u);
                while (k0 >= piece_end(a, pi)) pi++;
                p = get_piece(a, pi);
            }
            const bool fast = k0 + d.nsamp <= p.k_end;
            if (fast) {
                issue_next(s);
                refilled = true;
                if (p.tab == kNoTab)
                    stream_tile_direct<C, IN, OUT>(raw, out_s, p, k0, lane, win);   // the COLUMN window doubles as the plateau scratch
                else if (pi == a.smem_piece)
                    stream_tile<C, IN, OUT, kTabShared>(raw, out_s, p, k0, tab_s, plane_len, lane);
                else
                    stream_tile<C, IN, OUT, kTabGlobal>(raw, out_s, p, k0, a.tables + p.tab, 0, lane);
            } else {
                stream_tile_slow<C, IN, OUT>(a, pi, in_s, out_s, k0, d.nsamp, lane);
                __syncwarp();
            }
        }
        if (!refilled) issue_next(s);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(gout + (size_t)(d.k0 + d.skip) * kOutBps, out_s + d.skip * kOutBps, (d.nsamp - d.skip) * kOutBps);
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait_all();

    // sub-granule end of the buffer (fewer than kGran samples): pipeline 0 mixes it straight from
    // global memory, sample by sample
    if (a.tail_begin < a.nsamples && pipe == 0) {
        pi = find_piece(a, 0, a.tail_begin);
        p = get_piece(a, pi);
        for (uint32_t k = a.tail_begin + lane; k < a.nsamples; k += 32) {
            if (k >= p.k_end) {
                pi = find_piece(a, pi, k);
                p = get_piece(a, pi);
            }
            const float2 ph = phasor(p.r, piece_samplenum(p, k - p.k_begin));
            store_sample<OUT>(a.out, k, cmul_unfused(load_sample<IN>(a.in, k), ph));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pipelines of Q warps (round 2 experiment, NOT the product path: it lost the A/B, see doppler_b200.cu).  At 8 bytes per sample the segmented kernel is bound by instruction issue,
// and a third of what it issues is per-TILE work: barrier wait and tile loads, unit expansion and bulk-load issue, fence and
// bulk-store issue -- ~150 instructions per 512-sample tile (ncu source view, profiles/r02_ncu_cfg3_before.txt).  Here Q
// warps share one pipeline: one bulk load brings a tile Q times as large, every warp multiplies its own sub-tile exactly
// as before (same per-warp code), one bulk store takes the result; the expansion and both issues are paid once per Q
// sub-tiles, by the pipeline's first warp.  Shared memory per SM is unchanged (Q sub-tiles per stage instead of Q stages
// of one); the warps of a pipeline meet at a named barrier (bar.sync) where a single warp used __syncwarp.
template <int IN, int OUT, int WARPS_, int S_, int U_, int Q_>
struct PipeCfg {
    using W = StreamCfg<IN, OUT, WARPS_, S_, U_>;   // the per-warp view the compute code is written against
    static constexpr int WARPS = WARPS_, S = S_, U = U_, Q = Q_;
    static constexpr int G = W::G, kGran = W::kGran, kRow = W::kRow;
    static constexpr int kPipes = WARPS / Q;                          // pipelines per CTA
    static constexpr int kTileSamples = Q * W::kTileSamples;          // what the work decomposition (TileSrc, host walk) sees
    static constexpr int kTileIn = Q * W::kTileIn, kTileOut = Q * W::kTileOut;
    static constexpr int kRing = S * (kTileIn + kTileOut);            // per pipeline
    static constexpr int kBarBytes = ((kPipes * S * 8 + 127) / 128) * 128;
    static constexpr int kDescBytes = S * (int)sizeof(TileDesc);      // per pipeline
    static constexpr int kWinIters = (kTileSamples + kWinLead + 1 + Q * 32 - 1) / (Q * 32);   // per thread of the pipeline
    static constexpr int kWinPlane = ((Q * 32 * kWinIters / G + 15) / 16) * 16 + 16 / G;
    static constexpr int kWinBytes = G * kWinPlane * 8;               // per pipeline
    static constexpr int kFixedSmem = kBarBytes + kPipes * (kRing + kDescBytes + kWinBytes);
    static_assert(WARPS % Q == 0 && kPipes <= 15, "one named barrier per pipeline");
    static_assert(G * kWinPlane >= Q * W::kPlateauEntries, "the COLUMN window doubles as the warps' plateau scratch");
    __host__ __device__ static constexpr uint32_t plane_len(uint32_t period) { return W::plane_len(period); }
    __host__ __device__ static constexpr uint32_t table_bytes(uint32_t period) { return W::table_bytes(period); }
};

// the COLUMN window of a Q-warp pipeline, evaluated by all of its threads (tq = thread index within the pipeline)
template <typename P>
__device__ __forceinline__ void eval_window_pipe(float2* win, float r, uint32_t period, uint32_t phase0, uint32_t tq)
{
    constexpr int G = P::G, NW = P::kWinIters, PL = P::kWinPlane, NT = P::Q * 32;
    const uint32_t f0 = phase0 >= (uint32_t)kWinLead ? phase0 - kWinLead : phase0 + period - kWinLead;
    const bool mono = f0 + (uint32_t)(NT * NW) <= period;   // no period wrap inside the window
    db_window_t dt;
    const int range = mono ? classify_tile(r, f0 + 1u, f0 + (uint32_t)(NT * NW), dt) : (int)kRangeGeneric;
    float2* dst = win + (tq % G) * PL + tq / G;   // entry e = tq + NT*it -> plane e % G, slot e / G
    const uint32_t nl = f0 + 1u + tq;
    if (range == kRangeLarge) {
#pragma unroll 3
        for (int it = 0; it < NW; it++) dst[it * (NT / G)] = phasor_fast<kRangeLarge>(theta_of(r, nl + (uint32_t)(NT * it)), dt);
    } else if (range == kRangeMedium) {
#pragma unroll 3
        for (int it = 0; it < NW; it++) dst[it * (NT / G)] = phasor_fast<kRangeMedium>(theta_of(r, nl + (uint32_t)(NT * it)), dt);
    } else if (range == kRangeSmall) {
#pragma unroll 3
        for (int it = 0; it < NW; it++) dst[it * (NT / G)] = phasor_fast<kRangeSmall>(theta_of(r, nl + (uint32_t)(NT * it)), dt);
    } else if (range == kRangeTiny) {
#pragma unroll 3
        for (int it = 0; it < NW; it++) dst[it * (NT / G)] = phasor_fast<kRangeTiny>(theta_of(r, nl + (uint32_t)(NT * it)), dt);
    } else {
#pragma unroll 1
        for (int it = 0; it < NW; it++) {
            uint32_t f = f0 + tq + (uint32_t)(NT * it);
            if (f >= period) f -= period;
            if (f >= period) f %= period;
            dst[it * (NT / G)] = phasor(r, f + 1u);
        }
    }
}

template <int IN, int OUT, int WARPS, int S, int U, int Q>
__global__ void __launch_bounds__(WARPS * 32, 1) mix_pipe_kernel(const __grid_constant__ MixArgs a)
{
    using P = PipeCfg<IN, OUT, WARPS, S, U, Q>;
    using W = typename P::W;
    constexpr int G = W::G;
    constexpr uint32_t kInBps = IN == I16 ? 4 : 8, kOutBps = OUT == I16 ? 4 : 8;
    constexpr uint32_t kSub = W::kTileSamples;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* rings = smem + P::kBarBytes;
    unsigned char* descs = rings + P::kPipes * P::kRing;
    unsigned char* wins = descs + P::kPipes * P::kDescBytes;
    float2* tab_s = reinterpret_cast<float2*>(smem + P::kFixedSmem);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t pl = warp / Q, sub = warp % Q, tq = sub * 32 + lane;   // pipeline within the CTA, warp within the pipeline
    const bool issuer = tq == 0;
    uint64_t* full = bars + pl * S;
    unsigned char* ring_in = rings + pl * P::kRing;
    unsigned char* ring_out = ring_in + S * P::kTileIn;
    TileDesc* desc_ring = reinterpret_cast<TileDesc*>(descs + pl * P::kDescBytes);
    float2* win = reinterpret_cast<float2*>(wins + pl * P::kWinBytes);
    auto pipe_sync = [&] { asm volatile("bar.sync %0, %1;" ::"r"(1u + pl), "r"((uint32_t)(Q * 32)) : "memory"); };

    if (issuer) {
        for (int s = 0; s < S; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pipe_sync();

    const uint32_t pipe = blockIdx.x * P::kPipes + pl, npipes = gridDim.x * P::kPipes;
    const unsigned char* gin = static_cast<const unsigned char*>(a.in);
    unsigned char* gout = static_cast<unsigned char*>(a.out);

    // the pipeline's first warp expands the work units and issues the loads; the descriptor reaches the others through shared memory
    TileSrc<P> src;
    if (sub == 0) src.init(a, lane, npipes);
    auto issue_next = [&](uint32_t s) {   // all lanes of the pipeline's first warp
        if (sub != 0) return;
        TileDesc d;
        const bool more = src.fetch(a, lane, d);
        if (lane == 0) {
            if (more) {
                const uint32_t bytes = (d.nsamp - d.skip) * kInBps;
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(ring_in + s * P::kTileIn + d.skip * kInBps, gin + (size_t)(d.k0 + d.skip) * kInBps, bytes, &full[s]);
            } else {
                d.k0 = d.nsamp = d.seg = d.info = d.phase0 = d.period = d.skip = 0;
                d.r = 0.0f;
            }
            desc_ring[s] = d;
        }
    };
    for (uint32_t s = 0; s < (uint32_t)S; s++) issue_next(s);

    // (optionally) one piece's table de-interleaved into shared memory
    uint32_t plane_len = 0;
    if (a.smem_piece != kNoPiece) {
        const DevPiece sp = get_piece(a, a.smem_piece);
        plane_len = P::plane_len(sp.period);
        const float2* srct = a.tables + sp.tab;
        for (uint32_t e = threadIdx.x; e < sp.period + (uint32_t)P::kRow; e += WARPS * 32)
            tab_s[(e % G) * plane_len + e / G] = __ldg(srct + e % sp.period);
    }
    __syncthreads();

    uint32_t pi = 0;
    DevPiece p = get_piece(a, 0);
    for (uint32_t i = 0;; i++) {
        const uint32_t s = i % S;
        pipe_sync();   // desc[s] is visible; the previous tile's readers of the window / scratch are done
        const TileDesc d = desc_ring[s];
        if (d.nsamp == 0) break;
        const unsigned char* in_s = ring_in + s * P::kTileIn + sub * W::kTileIn;
        unsigned char* out_s = ring_out + s * P::kTileOut + sub * W::kTileOut;
        mbar_wait(&full[s], (i / S) & 1u);
        uint32_t raw[U][4];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned char* srcp = in_s + (u * 32 + lane) * W::kGroupIn;
            if constexpr (W::kGroupIn == 16) {
                const uint4 w = *reinterpret_cast<const uint4*>(srcp);
                raw[u][0] = w.x, raw[u][1] = w.y, raw[u][2] = w.z, raw[u][3] = w.w;
            } else {
                const uint2 w = *reinterpret_cast<const uint2*>(srcp);
                raw[u][0] = w.x, raw[u][1] = w.y, raw[u][2] = 0, raw[u][3] = 0;
            }
        }
        if (issuer) bulk_wait_read<S - 1>();   // the store that last read out[s] (tile i - S) has drained
        pipe_sync();                           // every warp holds its sub-tile and the descriptor in registers
        const uint32_t k0s = d.k0 + sub * kSub;   // first sample of this warp's sub-tile
        bool refilled = false;
        if (d.info & kColFlag) {
            issue_next(s);
            refilled = true;
            if (d.info & kColFirst) {
                eval_window_pipe<P>(win, d.r, d.period, d.phase0, tq);
                pipe_sync();
            }
            stream_tile_column<W, IN, OUT, P::kWinPlane>(raw, out_s, win + sub * (kSub / G), d.info & 0xffu, lane);
        } else {
            const uint32_t k0 = d.k0;
            if (k0 >= p.k_end || k0 < p.k_begin) {
                pi = max(d.phase0, k0 >= p.k_end ? pi : 0u);
                while (k0 >= piece_end(a, pi)) pi++;
                p = get_piece(a, pi);
            }
            const bool fast = k0 + d.nsamp <= p.k_end;   // the whole pipeline tile inside one piece
            if (fast) {
                issue_next(s);
                refilled = true;
                if (p.tab == kNoTab)
                    stream_tile_direct<W, IN, OUT>(raw, out_s, p, k0s, lane, win + sub * W::kPlateauEntries);
                else if (pi == a.smem_piece)
                    stream_tile<W, IN, OUT, kTabShared>(raw, out_s, p, k0s, tab_s, plane_len, lane);
                else
                    stream_tile<W, IN, OUT, kTabGlobal>(raw, out_s, p, k0s, a.tables + p.tab, 0, lane);
            } else {
                const uint32_t done_before = sub * kSub;
                const uint32_t nsub = d.nsamp > done_before ? (d.nsamp - done_before < kSub ? d.nsamp - done_before : kSub) : 0u;
                stream_tile_slow<W, IN, OUT>(a, pi, in_s, out_s, k0s, nsub, lane);
                pipe_sync();   // the slow path reads in[s] itself: refill only when every warp is through
            }
        }
        if (!refilled) issue_next(s);
        fence_async_smem();
        pipe_sync();
        if (issuer) {
            bulk_s2g(gout + (size_t)(d.k0 + d.skip) * kOutBps, ring_out + s * P::kTileOut + d.skip * kOutBps, (d.nsamp - d.skip) * kOutBps);
            bulk_commit();
        }
    }
    if (issuer) bulk_wait_all();

    // sub-granule end of the buffer (fewer than kGran samples): pipeline 0's first warp mixes it straight from global memory
    if (a.tail_begin < a.nsamples && pipe == 0 && sub == 0) {
        pi = find_piece(a, 0, a.tail_begin);
        p = get_piece(a, pi);
        for (uint32_t k = a.tail_begin + lane; k < a.nsamples; k += 32) {
            if (k >= p.k_end) {
                pi = find_piece(a, pi, k);
                p = get_piece(a, pi);
            }
            const float2 ph = phasor(p.r, piece_samplenum(p, k - p.k_begin));
            store_sample<OUT>(a.out, k, cmul_unfused(load_sample<IN>(a.in, k), ph));
        }
    }
}

// The lean variant for launches that are ONE GRID segment of whole tiles (const mode, the BASELINE
// metric): no segment table, no descriptors -- tile i of pipeline p is tile p + i * npipes.  About
// 100 fewer issued instructions per tile than the segmented loop above, which is what keeps the
// table-in-shared-memory path HBM-bound at 8 samples per lane per tile.  The ragged end (fewer samples
// than a tile) is mixed from global memory by the pipeline next in line.
template <int IN, int OUT, int WARPS, int S, int U>
__global__ void __launch_bounds__(WARPS * 32, 1) mix_grid_kernel(const __grid_constant__ MixArgs a)
{
    using C = StreamCfg<IN, OUT, WARPS, S, U>;
    constexpr int G = C::G;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* rings = smem + C::kBarBytes;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // launches of the direct-evaluation shape carry a per-warp plateau scratch between the rings and the table
    float2* scratch = a.plateau_scratch ? reinterpret_cast<float2*>(smem + C::kGridSmem + warp * C::kPlateauBytes) : nullptr;
    float2* tab_s = reinterpret_cast<float2*>(smem + C::kGridSmem + (a.plateau_scratch ? WARPS * C::kPlateauBytes : 0));

    uint64_t* full = bars + warp * S;
    unsigned char* ring_in = rings + warp * C::kRing;
    unsigned char* ring_out = ring_in + S * C::kTileIn;

    // start-up: every warp owns its barriers (lane 0 initialises them, issues on them, the warp waits on them),
    // so the first loads go out before the CTA stages the table: their latency hides the staging
    if (lane == 0) {
        for (int s = 0; s < S; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const uint32_t pipe = blockIdx.x * WARPS + warp, npipes = gridDim.x * WARPS;
    const uint32_t ntiles = a.nunits;   // one GRID segment of whole tiles: unit = tile
    const uint32_t mine = pipe < ntiles ? (ntiles - pipe + npipes - 1) / npipes : 0;
    const unsigned char* gin = static_cast<const unsigned char*>(a.in);
    unsigned char* gout = static_cast<unsigned char*>(a.out);

    auto issue_load = [&](uint32_t i) {   // lane 0 only
        const uint32_t s = i % S;
        mbar_expect_tx(&full[s], C::kTileIn);
        bulk_g2s(ring_in + s * C::kTileIn, gin + (size_t)(pipe + i * npipes) * C::kTileIn, C::kTileIn, &full[s]);
    };
    if (lane == 0)
        for (uint32_t i = 0; i < (uint32_t)S && i < mine; i++) issue_load(i);

    // (optionally) one piece's table de-interleaved into shared memory
    uint32_t plane_len = 0;
    if (a.smem_piece != kNoPiece) {
        const DevPiece sp = get_piece(a, a.smem_piece);
        plane_len = C::plane_len(sp.period);
        const float2* src = a.tables + sp.tab;
        for (uint32_t e = threadIdx.x; e < sp.period + (uint32_t)C::kRow; e += WARPS * 32)
            tab_s[(e % G) * plane_len + e / G] = __ldg(src + e % sp.period);   // replicated past the period: no wrap in a row
    }
    __syncthreads();

    uint32_t pi = 0;
    DevPiece p = get_piece(a, 0);
    for (uint32_t i = 0; i < mine; i++) {
        const uint32_t s = i % S;
        const uint32_t tile = pipe + i * npipes;
        const uint32_t k0 = tile * C::kTileSamples;
        const unsigned char* in_s = ring_in + s * C::kTileIn;
        unsigned char* out_s = ring_out + s * C::kTileOut;
        if (k0 >= p.k_end) {
            pi = find_piece(a, pi, k0);
            p = get_piece(a, pi);
        }
        const bool fast = k0 + C::kTileSamples <= p.k_end;
        mbar_wait(&full[s], (i / S) & 1u);
        if (fast) {
            uint32_t raw[U][4];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const unsigned char* src = in_s + (u * 32 + lane) * C::kGroupIn;
                if constexpr (C::kGroupIn == 16) {
                    const uint4 w = *reinterpret_cast<const uint4*>(src);
                    raw[u][0] = w.x, raw[u][1] = w.y, raw[u][2] = w.z, raw[u][3] = w.w;
                } else {
                    const uint2 w = *reinterpret_cast<const uint2*>(src);
                    raw[u][0] = w.x, raw[u][1] = w.y, raw[u][2] = 0, raw[u][3] = 0;
                }
            }
            if (lane == 0) bulk_wait_read<S - 1>();   // the store that last read out[s] (tile i - S) has drained
            __syncwarp();
            if (lane == 0 && i + S < mine) issue_load(i + S);   // in[s] is in registers now: refill it
            if (p.tab == kNoTab)
                stream_tile_direct<C, IN, OUT>(raw, out_s, p, k0, lane, scratch);
            else if (pi == a.smem_piece)
                stream_tile<C, IN, OUT, kTabShared>(raw, out_s, p, k0, tab_s, plane_len, lane);
            else
                stream_tile<C, IN, OUT, kTabGlobal>(raw, out_s, p, k0, a.tables + p.tab, 0, lane);
        } else {
            if (lane == 0) bulk_wait_read<S - 1>();
            __syncwarp();
            stream_tile_slow<C, IN, OUT>(a, pi, in_s, out_s, k0, C::kTileSamples, lane);
            __syncwarp();
            if (lane == 0 && i + S < mine) issue_load(i + S);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(gout + (size_t)tile * C::kTileOut, out_s, C::kTileOut);
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait_all();

    // ragged end of the buffer (fewer samples than a tile; may not be 16-byte granular): the
    // pipeline next in line mixes it straight from global memory, sample by sample
    const uint32_t tail0 = ntiles * C::kTileSamples;
    if (tail0 < a.nsamples && pipe == ntiles % npipes) {
        pi = find_piece(a, 0, tail0);
        p = get_piece(a, pi);
        for (uint32_t k = tail0 + lane; k < a.nsamples; k += 32) {
            if (k >= p.k_end) {
                pi = find_piece(a, pi, k);
                p = get_piece(a, pi);
            }
            const float2 ph = phasor(p.r, piece_samplenum(p, k - p.k_begin));
            store_sample<OUT>(a.out, k, cmul_unfused(load_sample<IN>(a.in, k), ph));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Small launches.  One second of the reference's own operating point (rtl_fm at 1.024 Msps, README.md:53) is
// 4 MB; one pump block (main.rs:49) is 8 KB.  At these sizes the persistent bulk-async kernels above spend their
// time on start-up (200 KB shared-memory carve-out, table staging, CTA barrier, pipeline fill and drain: ~11 us
// per launch, profiles/r01_sweep_final2.md) while HBM needs well under a microsecond.  This kernel is the
// latency-shaped variant: grid sized to the input, no shared memory, one 16-byte group per thread per step
// (LDG -> registers -> STG), the phasor table of a short period read through L1 (it is a few KB and stays there),
// everything else evaluated per sample with the generic routine.  It also serves the per-block host path
// zero-copy: `in` / `out` may be mapped pinned HOST memory, and `done` lets the host wait on a flag in its own
// memory instead of a stream synchronisation.
constexpr int kSmallThreads = 256;

struct SmallDone {
    uint32_t* counter;          // device: CTAs finished (returns to 0)
    volatile uint32_t* flag;    // mapped host memory: receives `token` when every CTA's stores are visible system-wide
    uint32_t token;
};

// V groups per thread per step, all V loads issued before the first is used (memory-level parallelism: the tiny
// zero-copy launches are one CTA and pay a PCIe round trip per dependent load; the mid sizes want bytes in flight).
// CTA `cta` of `nctas` takes an equal, contiguous share of the groups.
// COHERENT (the resident kernel below, which outlives many calls): input is read with ld.global.cv and tables with
// ld.global.cg, so that neither a previous call's block in the same host buffer nor a neighbouring table built after the
// kernel started can be served from this SM's L1.
// 16 bytes of flagged units: {w0, flag, w1, flag}.  Each 8-byte half is delivered whole (what NCCL's LL protocol rests on).
__device__ __forceinline__ void store_units2(void* units, uint32_t first_unit, uint32_t w0, uint32_t w1, uint32_t flag)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<uint2*>(units) + first_unit), "r"(w0), "r"(flag), "r"(w1),
                 "r"(flag)
                 : "memory");
}
__device__ __forceinline__ void store_unit(void* units, uint32_t unit, uint32_t w, uint32_t flag)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(units) + unit), "r"(w), "r"(flag) : "memory");
}

// COHERENT also changes where the result goes: a.out is an array of 8-byte UNITS, unit u = {word u of the result, `flag`}
// (the resident kernel's fence-free hand-over: the host takes a word when its flag shows the request number).
// STAGE (resident kernel, 16-byte groups): the input does not travel into registers but, by cp.async, into `stage` (shared memory,
// 16 bytes per group of the block).  A block without a table evaluates its phasors with the generic routine -- calls -- and
// around a call the compiler parks the registers that loads are in flight to, which waits for the loads: the ~2.4 us of
// evaluation ran after the round trip to host memory instead of under it.  With the input on its way to shared memory no
// register is pending: 9.6 -> 8.7 us per table-less block (10.2 -> 9.0 paced) -- the case of realtime track mode, a new ratio
// for every block, and of any shift whose period is beyond a table.  Blocks whose pieces all have tables pay 0.3 us for the
// detour (7.17 -> 7.44 on the same box); choosing per block at run time was measured and is worse for both (7.77 / 9.54: both
// forms in one body cost registers and scheduling).  profiles/r02_percall_mailbox_ab.txt
template <int IN, int OUT, int V, bool COHERENT>
__device__ __forceinline__ void small_body(const MixArgs& a, uint32_t cta, uint32_t nctas, uint32_t flag = 0, unsigned char* stage = nullptr)
{
    constexpr int G = group_samples(IN, OUT);
    constexpr bool kStaged = COHERENT && !(IN == I16 && G == 2);   // (cp.async past L1 moves 16 bytes; the 8-byte groups stay in registers)
    constexpr uint32_t kStep = kSmallThreads * V;
    const uint32_t ngroups = a.nsamples / G;
    uint32_t pi = 0;
    DevPiece p = get_piece(a, 0);
    // every CTA takes an equal, contiguous share of the groups (a grid-stride loop over fixed steps leaves whole steps
    // unevenly spread: 2.06 steps per CTA means a third of the CTAs run 3 while the rest wait)
    const uint32_t g_begin = (uint32_t)((uint64_t)ngroups * cta / nctas);
    const uint32_t g_end = (uint32_t)((uint64_t)ngroups * (cta + 1) / nctas);
    // ragged end (fewer samples than a group): one thread each
    auto tail_index = [&]() { return ngroups * G + cta * kSmallThreads + threadIdx.x; };
    auto load_tail = [&](uint32_t tail) -> float2 {
        if constexpr (COHERENT) {
            if constexpr (IN == I16)
                return ingest_i16(__ldcv(reinterpret_cast<const uint32_t*>(a.in) + tail));
            else
                return __ldcv(reinterpret_cast<const float2*>(a.in) + tail);
        } else {
            return load_sample<IN>(a.in, tail);
        }
    };
    auto tail_phasor = [&](uint32_t tail) -> float2 {
        const DevPiece q = get_piece(a, find_piece(a, 0, tail));
        const uint32_t n = piece_samplenum(q, tail - q.k_begin);
        if (q.tab != kNoTab) {   // the piece's table holds the same bits, and reading it is no call with the groups' loads in flight
            const float2* e = a.tables + q.tab + (n - 1u);
            return COHERENT ? __ldcg(e) : __ldg(e);
        }
        return phasor(q.r, n);
    };
    float2 tail_smp = make_float2(0.f, 0.f), tail_ph = make_float2(0.f, 0.f);
    bool tail_ready = false;
    for (uint32_t base = g_begin; base < g_end; base += kStep) {
        uint32_t w[V][4];
#pragma unroll
        for (int v = 0; v < V; v++) {
            const uint32_t g = base + v * kSmallThreads + threadIdx.x;
            w[v][0] = w[v][1] = w[v][2] = w[v][3] = 0;
            if constexpr (kStaged) {
                if (g < g_end)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(stage + 16u * (g - g_begin))),
                                 "l"(reinterpret_cast<const uint4*>(a.in) + g)
                                 : "memory");
            } else if (g < g_end) {
                if constexpr (IN == I16 && G == 2) {
                    const uint2* src = reinterpret_cast<const uint2*>(a.in) + g;
                    const uint2 x = COHERENT ? __ldcv(src) : __ldcs(src);
                    w[v][0] = x.x, w[v][1] = x.y;
                } else {
                    const uint4* src = reinterpret_cast<const uint4*>(a.in) + g;
                    const uint4 x = COHERENT ? __ldcv(src) : __ldcs(src);
                    w[v][0] = x.x, w[v][1] = x.y, w[v][2] = x.z, w[v][3] = x.w;
                }
            }
        }
        if constexpr (kStaged) asm volatile("cp.async.commit_group;" ::: "memory");
        // the G phasors of group g: table entries, or direct evaluation where there is no table / a piece ends inside the group
        auto phasors = [&](uint32_t g, float2 (&out)[G]) {
            const uint32_t k0 = g * G;
            if (k0 >= p.k_end) {
                pi = find_piece(a, pi, k0);
                p = get_piece(a, pi);
            }
            if (k0 + G <= p.k_end && p.tab != kNoTab) {
                // inside one tabled piece: entries j .. j+G-1 of its table (replicated kTabPad >= G-1 entries past the period)
                const float2* tab = a.tables + p.tab + (piece_samplenum(p, k0 - p.k_begin) - 1u);
#pragma unroll
                for (int i = 0; i < G; i++) out[i] = COHERENT ? __ldcg(tab + i) : __ldg(tab + i);
            } else {
                uint32_t qi = pi;
                DevPiece q = p;
#pragma unroll
                for (int i = 0; i < G; i++) {
                    const uint32_t k = k0 + (uint32_t)i;
                    if (k >= q.k_end) {
                        qi = find_piece(a, qi, k);
                        q = get_piece(a, qi);
                    }
                    out[i] = phasor(q.r, piece_samplenum(q, k - q.k_begin));
                }
            }
        };
        // The phasors do not depend on the input.  The resident kernel (one CTA, latency is all that counts) fetches them
        // while the input is still on its way from host memory, instead of after it -- left to the compiler's scheduling the
        // same source read 5.9 or 6.6 us per block from one build to the next.  The empty asm ties the input's first use to
        // the phasors: everything above it is issued before anything waits for the input.  (The launched kernels keep the
        // fused form: the extra live registers would cost them a resident CTA per SM.)
        float2 ph[COHERENT ? V : 1][G];
        if constexpr (COHERENT) {
#pragma unroll
            for (int v = 0; v < V; v++) {
                const uint32_t g = base + v * kSmallThreads + threadIdx.x;
#pragma unroll
                for (int i = 0; i < G; i++) ph[v][i] = make_float2(0.f, 0.f);
                if (g < g_end) phasors(g, ph[v]);
            }
            // (the ragged end's load and phasor too, so that nothing of a block is issued after the input has arrived)
            if (!tail_ready && tail_index() < a.nsamples) {
                tail_smp = load_tail(tail_index());
                tail_ph = tail_phasor(tail_index());
                tail_ready = true;
            }
            if constexpr (kStaged) {
                asm volatile("cp.async.wait_all;" ::: "memory");   // this thread's copies have landed: it reads back only what it copied
#pragma unroll
                for (int v = 0; v < V; v++) {
                    const uint32_t g = base + v * kSmallThreads + threadIdx.x;
                    if (g < g_end) {
                        const uint4 x = *reinterpret_cast<const uint4*>(stage + 16u * (g - g_begin));
                        w[v][0] = x.x, w[v][1] = x.y, w[v][2] = x.z, w[v][3] = x.w;
                    }
                }
            }
#pragma unroll
            for (int v = 0; !kStaged && v < V; v++) {
                if constexpr (G == 4)
                    asm volatile("" : "+r"(w[v][0]), "+r"(w[v][1]), "+r"(w[v][2]), "+r"(w[v][3])
                                 : "f"(ph[v][0].x), "f"(ph[v][0].y), "f"(ph[v][1].x), "f"(ph[v][1].y), "f"(ph[v][2].x), "f"(ph[v][2].y), "f"(ph[v][3].x),
                                   "f"(ph[v][3].y));
                else
                    asm volatile("" : "+r"(w[v][0]), "+r"(w[v][1]), "+r"(w[v][2]), "+r"(w[v][3]) : "f"(ph[v][0].x), "f"(ph[v][0].y), "f"(ph[v][1].x), "f"(ph[v][1].y));
            }
        }
#pragma unroll
        for (int v = 0; v < V; v++) {
            const uint32_t g = base + v * kSmallThreads + threadIdx.x;
            if (g >= g_end) break;
            float2 smp[G], res[G];
            if constexpr (IN == I16) {
#pragma unroll
                for (int i = 0; i < G; i++) smp[i] = ingest_i16(w[v][i]);
            } else {
                smp[0] = make_float2(__uint_as_float(w[v][0]), __uint_as_float(w[v][1]));
                smp[1] = make_float2(__uint_as_float(w[v][2]), __uint_as_float(w[v][3]));
            }
            if constexpr (!COHERENT) phasors(g, ph[0]);
#pragma unroll
            for (int i = 0; i < G; i++) res[i] = cmul_unfused(smp[i], ph[COHERENT ? v : 0][i]);
            if constexpr (COHERENT) {
                if constexpr (OUT == I16 && G == 4) {
                    store_units2(a.out, 4u * g, egress_i16(res[0]), egress_i16(res[1]), flag);
                    store_units2(a.out, 4u * g + 2u, egress_i16(res[2]), egress_i16(res[3]), flag);
                } else if constexpr (OUT == I16) {
                    store_units2(a.out, 2u * g, egress_i16(res[0]), egress_i16(res[1]), flag);
                } else {
                    store_units2(a.out, 4u * g, __float_as_uint(res[0].x), __float_as_uint(res[0].y), flag);
                    store_units2(a.out, 4u * g + 2u, __float_as_uint(res[1].x), __float_as_uint(res[1].y), flag);
                }
            } else if constexpr (OUT == I16 && G == 4) {
                __stcs(reinterpret_cast<uint4*>(a.out) + g, make_uint4(egress_i16(res[0]), egress_i16(res[1]), egress_i16(res[2]), egress_i16(res[3])));
            } else if constexpr (OUT == I16) {
                __stcs(reinterpret_cast<uint2*>(a.out) + g, make_uint2(egress_i16(res[0]), egress_i16(res[1])));
            } else {
                __stcs(reinterpret_cast<float4*>(a.out) + g, make_float4(res[0].x, res[0].y, res[1].x, res[1].y));
            }
        }
    }
    const uint32_t tail = tail_index();
    if (tail < a.nsamples) {
        if (!tail_ready) {
            tail_smp = load_tail(tail);
            tail_ph = tail_phasor(tail);
        }
        const float2 res = cmul_unfused(tail_smp, tail_ph);
        if constexpr (!COHERENT)
            store_sample<OUT>(a.out, tail, res);
        else if constexpr (OUT == I16)
            store_unit(a.out, tail, egress_i16(res), flag);
        else
            store_units2(a.out, 2u * tail, __float_as_uint(res.x), __float_as_uint(res.y), flag);
    }
}

template <int IN, int OUT, int V>
__global__ void __launch_bounds__(kSmallThreads) mix_small_kernel(const __grid_constant__ MixArgs a, const SmallDone done)
{
    small_body<IN, OUT, V, false>(a, blockIdx.x, gridDim.x);
    if (done.flag) {
        __threadfence_system();   // this thread's stores (possibly into host memory) are visible before the flag can be
        __syncthreads();
        if (threadIdx.x == 0) {
            if (gridDim.x == 1) {
                *done.flag = done.token;
            } else if (atomicAdd(done.counter, 1u) == gridDim.x - 1u) {
                *done.counter = 0u;
                __threadfence_system();
                *done.flag = done.token;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Resident kernel: the per-block host path without a launch per block.
//
// The reference calls its mixer once per 8192-byte pump block (main.rs:49,70), and a realtime stream delivers such a block
// every few milliseconds.  Even the zero-copy small launch above spends most of its ~14 us in the launch itself (driver call,
// launch latency, kernel start-up).  So the first per-block call starts ONE CTA that stays on the chip and serves the following
// blocks, everything travelling through mapped pinned host memory:
//     host:   copy the block into the pinned staging buffer, plan it, write request n into the mailbox
//     device: warp 0 looks at the mailbox (one 128-byte read per look); on a new request the CTA mixes the block from the
//             staging buffer and writes every 32-bit word of the result next to the request number
//     host:   collects the result: a word is there when its neighbour shows n
// What a block costs is PCIe round trips (~2 us each), so the protocol is built to need few of them:
//   * A request is 128 bytes: four 32-byte SECTORS of seven payload words and a tag, the request number -- head word (types,
//     number of pieces), sample count, up to kRtPieces pieces of 8 words.  The buffers' addresses travel with the launch, not
//     with the request.  A sector is the unit the memory system fetches, so a read of one is a coherent snapshot; the host
//     writes a sector's payload before its tag (x86 stores are ordered), so a sector whose tag is n carries request n, and the
//     CTA takes the request when all four tags show the same new number.  (An earlier version tagged 64-byte lines and added
//     the sum of the words against a line being fetched as two sectors at different times; a sum does not tell a block of
//     2048 samples at base 100 from one of 2047 at base 101.  Tags per sector need no such argument and cost nothing: the
//     layouts interleaved on one box read 5.84-5.94 us per block either way, profiles/r02_percall_mailbox_ab.txt.)
//   * One look at a time, by warp 0 alone while the other warps wait at a barrier.  (Keeping four looks in flight, issued a
//     fraction of a microsecond apart so that a request need not wait for the previous look to return, was measured and LOST
//     3.4 us per block: reads of host memory still in flight delay the block's own reads.  With blocks sent back to back the
//     first look after a block always comes too early -- the result is still on its way to the host -- and costs a round trip;
//     delaying it by 1.2 us gained 0.2 us per block, by less it lost up to 0.5: not kept.  profiles/r02_percall_mailbox_ab.txt)
//   * The result needs no fence and no completion flag.  An 8-byte store is delivered whole, so the unit {word, n} is its own
//     completion signal (the scheme of NCCL's LL protocol, which runs over PCIe for the same reason); a system-scope fence before
//     a flag costs a round trip of ~3 us on top of the 16 KB this writes instead of 8.
// The kernel leaves by itself after `idle_ns` without a request (alive = 0, then one last look issued after that), and at once
// on a quit request; the host starts another when it finds alive == 0.  Only plans of up to kRtPieces pieces come here.
constexpr int kRtStageBytes = 32 << 10;                              // the largest block the resident kernel serves (doppler_b200.cu: kTinyStageBytes)
constexpr int kRtPieces = 3;
constexpr int kRtPieceWords = 8;                                     // DevPiece up to and including `shift` (small_body reads no more)
constexpr int kRtPayloadWords = 2 + kRtPieces * kRtPieceWords;       // head, nsamples, pieces
constexpr int kRtSectors = 4;                                        // of 7 payload words + tag
constexpr uint32_t kRtQuit = 0x80000000u;                            // head word: leave now
constexpr uint32_t kRtLast = 0x40000000u;                            // (device only) this request came with the last look
struct RtSector {
    uint32_t w[7];
    uint32_t tag;
};
struct RtMailbox {
    RtSector req[kRtSectors];   // host -> device
    volatile uint32_t alive;    // generation of the resident kernel (the host writes it before the launch); the kernel clears it when it leaves
    uint32_t pad0[15];
    volatile uint32_t served;   // device -> its successor: the latest request taken (a kernel that starts behind one that was still
                                // leaving must not take that one's last request again); the host never reads it (its own line)
    uint32_t pad1[15];
};
static_assert(sizeof(RtSector) == 32 && kRtSectors * 8 == 32, "the mailbox is one warp-wide read");
static_assert(kRtPayloadWords <= kRtSectors * 7, "the request fits its sectors");
static_assert(kRtPieces <= kInlinePieces && kRtPieceWords * 4 <= (int)offsetof(DevPiece, step_u), "pieces travel up to `shift`");

// head word of a request
__host__ __device__ constexpr uint32_t rt_head(int intype, int outtype, uint32_t npieces) { return (uint32_t)outtype | ((uint32_t)intype << 1) | (npieces << 2); }

__device__ __forceinline__ uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One look's verdict, computed by the whole warp from the word each lane read: the request number if the four sectors hold one
// whole request other than `last`, else `last`.
__device__ __forceinline__ uint32_t rt_verdict(uint32_t v, uint32_t last)
{
    const uint32_t t0 = __shfl_sync(0xffffffffu, v, 7), t1 = __shfl_sync(0xffffffffu, v, 15);
    const uint32_t t2 = __shfl_sync(0xffffffffu, v, 23), t3 = __shfl_sync(0xffffffffu, v, 31);
    return (t0 == t1 && t0 == t2 && t0 == t3) ? t0 : last;
}

static __global__ void __launch_bounds__(kSmallThreads, 1)
mix_resident_kernel(RtMailbox* mb, const void* in, void* out_units, const float2* tables, uint64_t idle_ns, uint32_t gen, uint32_t last)
{
    __shared__ MixArgs s_args;
    __shared__ uint32_t s_head, s_seq;
    __shared__ __align__(16) unsigned char s_stage[kRtStageBytes];   // the block's input on its way in (small_body: STAGE)
    const uint32_t tid = threadIdx.x;
    const volatile uint32_t* words = reinterpret_cast<const volatile uint32_t*>(mb);
    {
        const uint32_t s = mb->served;   // (written by a predecessor on the same stream: complete before this kernel started)
        if (s != 0 && (int32_t)(s - last) > 0) last = s;
    }
    for (uint32_t i = tid; i < sizeof(MixArgs) / 4; i += kSmallThreads) reinterpret_cast<uint32_t*>(&s_args)[i] = 0;
    __syncthreads();
    if (tid == 0) {
        s_args.in = in;
        s_args.out = out_units;
        s_args.tables = tables;
        s_args.smem_piece = 0xffffffffu;
        s_args.max_claim = 1;
    }
    for (;;) {
        if (tid < 32u) {
            // warp 0 looks for the next request; the other warps wait at the barrier below
            uint32_t v = 0, seq = last;
            bool leave = false;
            const uint64_t t0 = global_timer_ns();
            for (;;) {   // one look at a time (see above)
                v = words[tid];
                seq = rt_verdict(v, last);
                if (seq != last) break;
                if (__shfl_sync(0xffffffffu, global_timer_ns() - t0 > idle_ns ? 1 : 0, 0)) {   // (lane 0's clock: one verdict for the warp)
                    leave = true;
                    break;
                }
            }
            if (leave) {
                // idle: announce the departure, then look once more -- a request that raced with the time-out is still served
                if (tid == 0) {
                    if (mb->alive == gen) mb->alive = 0;
                    __threadfence_system();
                }
                __syncwarp();
                v = words[tid];
                seq = rt_verdict(v, last);
            }
            if (seq != last) {
                const uint32_t w = (tid >> 3) * 7u + (tid & 7u);   // payload index of this lane's word (tags aside)
                if ((tid & 7u) != 7u) {
                    if (w == 0) {
                        s_head = leave ? (v | kRtLast) : v;
                        s_seq = seq;
                        s_args.npieces = (v >> 2) & 7u;
                    } else if (w == 1) {
                        s_args.nsamples = v;
                    } else if (w < (uint32_t)kRtPayloadWords) {
                        const uint32_t pw = w - 2u;
                        reinterpret_cast<uint32_t*>(&s_args.inl[pw / kRtPieceWords])[pw % kRtPieceWords] = v;
                    }
                }
            } else if (tid == 0) {
                s_head = kRtQuit;   // nothing came with the last look
                s_seq = last;
            }
        }
        __syncthreads();
        const uint32_t head = s_head;
        if (head & kRtQuit) break;
        last = s_seq;
        if (tid == 0) mb->served = last;
        switch (head & 3u) {
        case 0: small_body<I16, I16, 4, true>(s_args, 0, 1, last, s_stage); break;
        case 1: small_body<I16, F32, 4, true>(s_args, 0, 1, last, s_stage); break;
        case 2: small_body<F32, I16, 4, true>(s_args, 0, 1, last, s_stage); break;
        default: small_body<F32, F32, 4, true>(s_args, 0, 1, last, s_stage); break;
        }
        if (head & kRtLast) break;   // (served after the departure was announced: the host will start a successor)
        __syncthreads();             // s_args / s_head are rewritten by the next request
    }
    if (tid == 0) {
        if (mb->alive == gen) mb->alive = 0;   // (a successor the host has already announced keeps its own mark)
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------
// Fused downstream stage (not in the reference): mix, then a decimate-by-M FIR, in one pass.
//   y[k] = the mixer's Complex<f32> result for stream sample k (exactly the arithmetic above), y[k] = 0 before the stream
//   z[m] = sum over t = 0 .. ntaps-1, in that order, of h[t] * y[m*M - t], one fused multiply-add per step (fmaf), re / im apart
// (specification: the definition in include/doppler_b200.h, restated on the CPU by the test checker).  The output is 1/M of the stream, so for host callers the
// device -> host copy -- half of the PCIe traffic of an i16 -> i16 mix -- shrinks by M.  A CTA stages the mixed samples its
// outputs need in shared memory (for even M skewed by one slot per M samples, so that the lanes' reads at stride M fall in
// distinct banks), then every thread accumulates one output.
struct DecimArgs {
    MixArgs mix;             // in, pieces, tables of this call (out unused)
    void* out;               // decimated output (outtype)
    const float2* hist;      // the ntaps-1 mixed samples before this call's first sample, oldest first
    float2* hist_next;       // (history pass) the ntaps-1 mixed samples before the NEXT call's first sample
    const float* taps;
    uint32_t ntaps, M;
    uint32_t first_out;      // call-relative index of the first sample whose stream position is a multiple of M
    uint32_t nout;           // outputs of this call
    uint32_t out_per_cta;    // outputs per CTA step
    uint32_t skew;           // 1: staged sample j lives at slot j + j / M
};
constexpr int kDecimThreads = 256;

// the mixer's result for call-relative sample i (any piece, table or direct evaluation)
template <int IN>
__device__ __forceinline__ float2 mix_one(const MixArgs& a, uint32_t i, uint32_t& pi, DevPiece& p)
{
    if (i >= p.k_end || i < p.k_begin) {
        pi = find_piece(a, i < p.k_begin ? 0u : pi, i);
        p = get_piece(a, pi);
    }
    const uint32_t n = piece_samplenum(p, i - p.k_begin);
    const float2 ph = p.tab != kNoTab ? __ldg(a.tables + p.tab + (n - 1u)) : phasor(p.r, n);
    return cmul_unfused(load_sample<IN>(a.in, i), ph);
}

template <int IN, int OUT>
__global__ void __launch_bounds__(kDecimThreads) mix_decimate_kernel(const __grid_constant__ DecimArgs d)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float* taps_s = reinterpret_cast<float*>(smem);
    float2* y_s = reinterpret_cast<float2*>(smem + ((d.ntaps * 4 + 15) & ~15u));
    const uint32_t nh = d.ntaps - 1u, M = d.M, OT = d.out_per_cta;
    for (uint32_t t = threadIdx.x; t < d.ntaps; t += kDecimThreads) taps_s[t] = d.taps[t];
    uint32_t pi = 0;
    DevPiece p = get_piece(d.mix, 0);
    for (uint32_t ob = blockIdx.x * OT; ob < d.nout; ob += gridDim.x * OT) {
        const uint32_t ocount = d.nout - ob < OT ? d.nout - ob : OT;
        // staged slot j holds call-relative sample i0 + j, i0 = first_out + ob*M - nh (may be negative: history)
        const int64_t i0 = (int64_t)d.first_out + (int64_t)ob * M - (int64_t)nh;
        const uint32_t count = (ocount - 1u) * M + d.ntaps;
        __syncthreads();   // the previous step's readers are done (and the taps are staged)
        for (uint32_t j = threadIdx.x; j < count; j += kDecimThreads) {
            const int64_t i = i0 + (int64_t)j;
            const float2 y = i < 0 ? d.hist[(int64_t)nh + i] : mix_one<IN>(d.mix, (uint32_t)i, pi, p);
            y_s[d.skew ? j + j / M : j] = y;
        }
        __syncthreads();
        if (threadIdx.x < ocount) {
            uint32_t j = nh + threadIdx.x * M;            // slot index (unskewed) of sample m*M
            uint32_t rem = j % M;
            uint32_t slot = d.skew ? j + j / M : j;
            float re = 0.0f, im = 0.0f;
            for (uint32_t t = 0; t < d.ntaps; t++) {
                const float2 y = y_s[slot];
                const float h = taps_s[t];
                re = __fmaf_rn(h, y.x, re);
                im = __fmaf_rn(h, y.y, im);
                // one sample back; crossing a multiple of M skips the skew slot
                if (rem == 0u) {
                    rem = M - 1u;
                    slot -= 1u + d.skew;
                } else {
                    rem -= 1u;
                    slot -= 1u;
                }
            }
            store_sample<OUT>(d.out, ob + threadIdx.x, make_float2(re, im));
        }
    }
}

// History pass: the last ntaps-1 mixed samples of the stream after this call (older ones come from the previous history).
template <int IN>
__global__ void __launch_bounds__(kDecimThreads) decim_history_kernel(const __grid_constant__ DecimArgs d)
{
    const uint32_t nh = d.ntaps - 1u;
    uint32_t pi = 0;
    DevPiece p = get_piece(d.mix, 0);
    for (uint32_t j = blockIdx.x * kDecimThreads + threadIdx.x; j < nh; j += gridDim.x * kDecimThreads) {
        const int64_t i = (int64_t)d.mix.nsamples - (int64_t)nh + (int64_t)j;
        d.hist_next[j] = i < 0 ? d.hist[(int64_t)nh + i] : mix_one<IN>(d.mix, (uint32_t)i, pi, p);
    }
}

// Phasor table of one shift: entry j (0 <= j < period + kTabPad) = phasor(r, (j mod period) + 1).
static __global__ void __launch_bounds__(kThreads) build_phasor_table_kernel(float2* tab, float r, uint32_t period, uint32_t entries)
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
static __global__ void phasor_probe_kernel(float r, uint32_t n0, uint32_t count, float* c, float* s)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float2 ph = phasor(r, n0 + i);
    c[i] = ph.x;
    s[i] = ph.y;
}

static __global__ void sincosf_probe_kernel(uint32_t first, uint32_t stride, uint32_t count, float* s, float* c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const db_sincos_t r = db_sincosf_glibc(__uint_as_float(first + i * stride));
    s[i] = r.s;
    c[i] = r.c;
}

}  // namespace dmix
