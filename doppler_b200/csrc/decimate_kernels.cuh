// decimate_kernels.cuh -- the fused mix + decimate-by-M FIR stage, register-blocked (not in the reference).
//
// Specification (include/doppler_b200.h, restated by oracle_mix_decimate): y[k] = the mixer's Complex<f32> result for stream
// sample k, z[m] = sum over t = 0 .. ntaps-1, IN THAT ORDER, of fmaf(h[t], y[m*M - t], acc), re / im apart.  The order is part
// of the contract, so every output is one serial chain of ntaps fused multiply-adds; the freedom left is which thread runs
// which chains and where h and y come from.
//
// What bounds the stage on B200 (8 B in + 4/M B out per sample for f32 -> i16: 760 Gsample/s at the copy peak, i.e. 3.3
// samples per SM per clock at 1.55 GHz):
//   * the load/store unit (128 B per clock per SM, shared by global loads, table lookups and shared memory): reading y once per
//     tap and output -- the first version, mix_decimate_kernel in mixer_kernels.cuh -- moves 8 * ntaps / M = 49 B of shared memory
//     per input sample; and a tap read per multiply-add doubles that even when it is a broadcast;
//   * instruction issue (128 lane-instructions per clock): ~94 issued per input sample in the first version.
// So here
//   * a thread owns R = 4 CONSECUTIVE outputs m0 .. m0+3 and walks the samples they need once, newest first: sample m0*M + e,
//     e = 3M down to -(ntaps-1), feeds output k with tap k*M - e wherever that is a tap.  One 8-byte shared-memory read serves up
//     to four multiply-adds: (3M + ntaps) / 4M reads per input sample instead of ntaps / M (2.3 instead of 6.1 at M = 8, 49 taps);
//   * the taps never touch the load/store unit: the host lays them out per walk position u = 3M - e as four (h, h) pairs
//     {h[u-3M], h[u-2M], h[u-M], h[u]} in the KERNEL PARAMETERS, u is warp-uniform, and the compiler reads them with uniform
//     constant loads (LDCU) straight into uniform registers that FFMA2 takes as an operand (packed re/im: one issue slot per
//     tap and output);
//   * the walk is cut into seven segments of constant active output range [klo, khi] (compile-time per filter shape, bounds from
//     the host), so the inner loops carry no predicates and issue exactly ntaps multiply-adds per output;
//   * lanes read y at a stride of 4M samples; one padding slot per 4M samples (slot(j) = j + j / 4M) makes the stride odd in
//     8-byte units, which is bank-conflict free.  The walk crosses a padding slot at warp-uniform positions.
// Phase A of a CTA step (mix the samples the step's outputs need into shared memory) runs 16-byte global loads, four in flight
// per thread, and table phasors through L1; phase B is the walk.  Several CTAs per SM overlap each other's phases.
//
// Launches outside the envelope (too many taps for the parameter block, M > 64, unaligned buffers) and pieces without a table
// keep the generic paths: mix_decimate_kernel, or the per-sample loop of phase A below.
#pragma once

#include "mixer_kernels.cuh"

namespace dmix {

constexpr int kDfR = 4;                 // outputs per thread
constexpr int kDfThreads = 256;
constexpr int kDfMaxTq = 224;           // walk positions (3M + ntaps) the parameter block holds
constexpr uint32_t kDfStageSlots = 8448;   // shared-memory slots (8 B) of one CTA step: 66 KB, three CTAs per SM

struct DecimFastArgs {
    DecimArgs d;
    uint32_t tb;          // threads of a CTA that own outputs (a multiple of 32): a CTA step makes 4 * tb outputs
    uint32_t lead;        // staged sample 0 is call-relative sample i0 - lead, so that it is a multiple of 4 (16-byte loads)
    uint32_t ctop;        // staged index, relative to a thread's base tid * 4M, of walk position 0: lead + ntaps - 1 + 3M
    uint32_t shape;       // min(3, (ntaps - 1) / M): selects the kernel instantiation
    uint32_t cuts[8];     // {0, M, 2M, 3M, ntaps, M + ntaps, 2M + ntaps, 3M + ntaps} sorted: segment i = [cuts[i], cuts[i+1])
    uint64_t tq[kDfMaxTq][kDfR];   // tq[u][k] = (h[t], h[t]) with t = u - (3 - k) * M, 0 outside the filter
};

__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// walk positions [u0, u1) with outputs KLO .. KHI active; p = this thread's slot of position u0, cm = how many more positions
// precede the next padding slot (warp-uniform)
template <int KLO, int KHI>
__device__ __forceinline__ void df_walk(const DecimFastArgs& A, const uint64_t*& p, uint32_t& cm, uint32_t rm, uint32_t u0, uint32_t u1,
                                        uint64_t (&acc)[kDfR])
{
    while (u0 < u1) {
        const uint32_t left = u1 - u0;
        const bool cross = cm < left;
        const uint32_t n = cross ? cm + 1u : left;
        const uint32_t ue = u0 + n;
#pragma unroll 4
        for (uint32_t u = u0; u < ue; ++u) {
            const uint64_t y = *p;
            --p;
#pragma unroll
            for (int k = KLO; k <= KHI; ++k) acc[k] = fma_f32x2(A.tq[u][k], y, acc[k]);
        }
        u0 = ue;
        if (cross) {
            --p;   // the padding slot
            cm = rm - 1u;
        } else {
            cm -= n;
        }
    }
}

// The walk's seven segments.  Output k is active at positions [(3 - k) * M, (3 - k) * M + ntaps); the eight bounds sorted
// are the segments' cuts (the host sorts them: DecimFastArgs::cuts), and which outputs are active between two neighbouring
// cuts depends only on SHAPE = min(3, (ntaps - 1) / M) -- so the ranges are compile-time and the dispatch is straight-line
// code (a switch on a run-time range compiles to an indexed branch, after which ptxas no longer keeps the tap reads on the
// uniform datapath).  Ties between cuts make empty segments.
template <int SHAPE, int I>
struct DfRange {
    //                                   segment:      0  1  2  3  4  5  6
    static constexpr int lo3[7] = {3, 2, 1, 0, 0, 0, 0}, hi3[7] = {3, 3, 3, 3, 2, 1, 0};   // ntaps > 3M
    static constexpr int lo2[7] = {3, 2, 1, 1, 0, 0, 0}, hi2[7] = {3, 3, 3, 2, 2, 1, 0};   // 2M < ntaps <= 3M
    static constexpr int lo1[7] = {3, 2, 2, 1, 1, 0, 0}, hi1[7] = {3, 3, 2, 2, 1, 1, 0};   // M < ntaps <= 2M
    static constexpr int lo0[7] = {3, 1, 2, 1, 1, 1, 0}, hi0[7] = {3, 0, 2, 0, 1, 0, 0};   // ntaps <= M (odd segments: gaps)
    static constexpr int lo = SHAPE == 3 ? lo3[I] : SHAPE == 2 ? lo2[I] : SHAPE == 1 ? lo1[I] : lo0[I];
    static constexpr int hi = SHAPE == 3 ? hi3[I] : SHAPE == 2 ? hi2[I] : SHAPE == 1 ? hi1[I] : hi0[I];
};

template <int SHAPE>
__device__ __forceinline__ void df_walk_all(const DecimFastArgs& A, const uint64_t*& p, uint32_t& cm, uint32_t rm, uint64_t (&acc)[kDfR])
{
    df_walk<DfRange<SHAPE, 0>::lo, DfRange<SHAPE, 0>::hi>(A, p, cm, rm, A.cuts[0], A.cuts[1], acc);
    df_walk<DfRange<SHAPE, 1>::lo, DfRange<SHAPE, 1>::hi>(A, p, cm, rm, A.cuts[1], A.cuts[2], acc);
    df_walk<DfRange<SHAPE, 2>::lo, DfRange<SHAPE, 2>::hi>(A, p, cm, rm, A.cuts[2], A.cuts[3], acc);
    df_walk<DfRange<SHAPE, 3>::lo, DfRange<SHAPE, 3>::hi>(A, p, cm, rm, A.cuts[3], A.cuts[4], acc);
    df_walk<DfRange<SHAPE, 4>::lo, DfRange<SHAPE, 4>::hi>(A, p, cm, rm, A.cuts[4], A.cuts[5], acc);
    df_walk<DfRange<SHAPE, 5>::lo, DfRange<SHAPE, 5>::hi>(A, p, cm, rm, A.cuts[5], A.cuts[6], acc);
    df_walk<DfRange<SHAPE, 6>::lo, DfRange<SHAPE, 6>::hi>(A, p, cm, rm, A.cuts[6], A.cuts[7], acc);
}

template <int IN>
__device__ __forceinline__ void df_mix_group(const uint4& raw, const float2* tab, uint32_t ph, float2* dst)
{
    if constexpr (IN == I16) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int s = 0; s < 4; s++) dst[s] = cmul_unfused(ingest_i16(w[s]), __ldg(tab + ph + s));
    } else {
        dst[0] = cmul_unfused(make_float2(__uint_as_float(raw.x), __uint_as_float(raw.y)), __ldg(tab + ph));
        dst[1] = cmul_unfused(make_float2(__uint_as_float(raw.z), __uint_as_float(raw.w)), __ldg(tab + ph + 1));
    }
}

template <int IN, int OUT, int SHAPE>
__global__ void __launch_bounds__(kDfThreads) mix_decimate_fast_kernel(const __grid_constant__ DecimFastArgs A)
{
    constexpr int R = kDfR, NT = kDfThreads;
    constexpr uint32_t V = IN == I16 ? 4 : 2;             // samples per 16-byte load
    constexpr uint32_t kInBps = IN == I16 ? 4 : 8;
    constexpr int UNR = 4;                                // loads in flight per thread
    extern __shared__ __align__(16) unsigned char smem[];
    float2* y_s = reinterpret_cast<float2*>(smem);
    const DecimArgs& d = A.d;
    const uint32_t M = d.M, RM = R * M, nh = d.ntaps - 1u, OT = A.tb * R, tid = threadIdx.x;
    const unsigned char* gin = static_cast<const unsigned char*>(d.mix.in);
    const uint32_t warp0 = __shfl_sync(0xffffffffu, tid & ~31u, 0);   // first thread of this warp, known uniform to the compiler
    uint32_t pi = 0;
    DevPiece p = get_piece(d.mix, 0);
    for (uint32_t ob = blockIdx.x * OT; ob < d.nout; ob += gridDim.x * OT) {
        const uint32_t ocount = d.nout - ob < OT ? d.nout - ob : OT;
        // staged index j holds call-relative sample ia + j (negative: history); ia is a multiple of 4
        const int64_t ia = (int64_t)d.first_out + (int64_t)ob * M - (int64_t)nh - (int64_t)A.lead;
        const uint32_t count = A.lead + (ocount - 1u) * M + d.ntaps;
        const uint32_t ngroups = (count + V - 1u) / V;
        __syncthreads();   // the previous step's walks are done
        // ---- phase A: mix the step's samples into shared memory
        bool fast = false;
        if (ia >= 0 && (uint64_t)ia + (uint64_t)ngroups * V <= (uint64_t)d.mix.nsamples) {
            const uint32_t first = (uint32_t)ia, last = first + ngroups * V;
            if (first >= p.k_end || first < p.k_begin) {
                pi = find_piece(d.mix, first < p.k_begin ? 0u : pi, first);
                p = get_piece(d.mix, pi);
            }
            fast = last <= p.k_end && p.tab != kNoTab;
        }
        if (fast) {
            const float2* tab = d.mix.tables + p.tab;
            const uint32_t period = p.period;
            // group g = tid + x * NT: sample ia + g * V, staged index j = g * V = jq * RM + jr, table phase ph
            uint32_t ph = piece_samplenum(p, (uint32_t)ia + tid * V - p.k_begin) - 1u;
            const uint32_t ph_step = (NT * V) % period;
            uint32_t jq = (tid * V) / RM, jr = (tid * V) % RM;
            const uint32_t jq_step = (NT * V) / RM, jr_step = (NT * V) % RM;
            const unsigned char* src = gin + ((size_t)ia + (size_t)tid * V) * kInBps;
            for (uint32_t g0 = tid; g0 < ngroups; g0 += UNR * NT) {
                uint4 raw[UNR];
#pragma unroll
                for (int x = 0; x < UNR; x++)
                    if (g0 + x * NT < ngroups) raw[x] = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)x * NT * 16));
                src += (size_t)UNR * NT * 16;
#pragma unroll
                for (int x = 0; x < UNR; x++) {
                    if (g0 + x * NT < ngroups) {
                        df_mix_group<IN>(raw[x], tab, ph, y_s + (jq * RM + jr) + jq);
                        ph += ph_step;
                        if (ph >= period) ph -= period;
                        jq += jq_step;
                        jr += jr_step;
                        if (jr >= RM) {
                            jr -= RM;
                            jq++;
                        }
                    }
                }
            }
        } else {
            for (uint32_t j = tid; j < count; j += NT) {
                const int64_t i = ia + (int64_t)j;
                float2 y = make_float2(0.0f, 0.0f);   // before the history / past the input: never read by a valid output
                if (i < 0) {
                    if (i + (int64_t)nh >= 0) y = d.hist[(int64_t)nh + i];
                } else if (i < (int64_t)d.mix.nsamples) {
                    y = mix_one<IN>(d.mix, (uint32_t)i, pi, p);
                }
                y_s[j + j / RM] = y;
            }
        }
        __syncthreads();
        // ---- phase B: thread tid owns outputs ob + 4 * tid + (0 .. 3)
        // (a warp-uniform condition: the walk's tap reads and loop control stay on the uniform datapath; lanes past the last
        // output walk staged slots inside the stage and store nothing)
        if (warp0 < A.tb && warp0 * R < ocount) {
            uint64_t acc[R];
#pragma unroll
            for (int k = 0; k < R; k++) acc[k] = 0ull;   // (+0.0f, +0.0f)
            const uint32_t c0 = A.ctop;
            const uint64_t* yp = reinterpret_cast<const uint64_t*>(y_s) + tid * (RM + 1u) + c0 + c0 / RM;
            uint32_t cm = c0 % RM;
            df_walk_all<SHAPE>(A, yp, cm, RM, acc);
            const uint32_t m0 = tid * R;
            float2 z[R];
#pragma unroll
            for (int k = 0; k < R; k++) unpack_f32x2(acc[k], z[k].x, z[k].y);
            if (m0 + R <= ocount) {
                if constexpr (OUT == I16) {
                    __stcs(reinterpret_cast<uint4*>(static_cast<uint32_t*>(d.out) + ob + m0),
                           make_uint4(egress_i16(z[0]), egress_i16(z[1]), egress_i16(z[2]), egress_i16(z[3])));
                } else {
                    float4* o = reinterpret_cast<float4*>(static_cast<float2*>(d.out) + ob + m0);
                    __stcs(o, make_float4(z[0].x, z[0].y, z[1].x, z[1].y));
                    __stcs(o + 1, make_float4(z[2].x, z[2].y, z[3].x, z[3].y));
                }
            } else {
#pragma unroll
                for (int k = 0; k < R; k++)
                    if (m0 + k < ocount) store_sample<OUT>(d.out, ob + m0 + k, z[k]);
            }
        }
    }
}

}  // namespace dmix
