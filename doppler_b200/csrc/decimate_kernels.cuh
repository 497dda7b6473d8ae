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
//     8-byte units, which is bank-conflict free.  Where the walk crosses a padding slot depends only on the launch, so the host
//     gives every walk position its slot offset next to its taps.
// Phase A of a CTA step (mix the samples the step's outputs need into shared memory) runs 16-byte global loads through two
// register buffers (the next batch is in flight while one is mixed; the first batch of the next step while this step walks)
// against a copy of the piece's phasor table in shared memory; phase B is the walk.  Several CTAs per SM overlap each other's
// phases.  CTAs are 256 threads, or 128 where a stage of 256 * 4 outputs would not fit (M > 8).
//
// Launches whose samples mostly lie in pieces WITHOUT a table (long periods, track mode) run in TWO passes: the mixer itself
// (mix_stream_kernel: COLUMN segments, range-specialised direct evaluation -- everything section 4.3/4.4 of DESIGN.md built for
// such pieces) writes the mixed stream as complex f32 behind the carried history in a scratch buffer, and this kernel runs
// over that buffer with input type C32 -- phase A is then a plain copy.  16 more bytes of HBM traffic per sample, but 2x faster
// than evaluating one generic sincosf per sample here.  Launches outside the envelope (too many taps for the parameter block,
// M > 64, unaligned buffers) keep the generic kernel, mix_decimate_kernel.
//
// Measured (B200, 256 M samples, tools/decim_bench.py, profiles/r02_decim_*): const f32 -> i16, M = 8, 49 taps 0.65 ms =
// 393 Gsample/s in = 0.52 of the stage's HBM roofline (first version: 1.78 ms, 0.19); i16 -> i16 0.26 of its 4.5 B/sample
// roofline; table-less launches (two passes) 1.18 ms, 0.28.  ncu: 52 issued instructions per input sample at 59 % issue
// utilisation, load/store-unit wavefronts at 66 % (two-thirds of them shared memory), 24 warps per SM -- the stage is bound by
// issue slots and the shared-memory pipe together, not by HBM (40 % of peak).  Restructurings that cut those counters but LOST
// on the clock, kept out of the tree: two lockstep output groups per thread (tap loads and loop control shared by eight
// chains: 39 instructions per sample, but half the warps per SM for the same shared memory: 0.75-0.80 ms) and one sample per
// lane in phase A (conflict-free shared memory, same result).  Shared memory per resident warp is what limits the stage.
#pragma once

#include "mixer_kernels.cuh"

namespace dmix {

constexpr int C32 = 2;                  // a third "input type" of this kernel: samples that are already mixed (complex f32), see below
constexpr int kDfR = 4;                 // outputs per thread
constexpr int kDfMaxThreads = 256;
constexpr int kDfMaxTq = 224;           // walk positions (3M + ntaps) the parameter block holds
constexpr uint32_t kDfStageSlots = 8448;   // shared-memory slots (8 B) of one CTA step by default: 66 KB, three CTAs per SM
constexpr uint32_t kDfTabCap = 1024;       // phasor-table entries (8 B) a CTA keeps in shared memory; longer tables are read through L1

struct alignas(16) DfTaps {
    uint64_t h[kDfR];   // h[k] = (tap, tap) of output k at this walk position: tap index u - (3 - k) * M, 0 outside the filter
    uint32_t off;       // byte offset, from the thread's base slot tid * (4M + 1), of the sample this position reads: (c + c / 4M) * 8
    uint32_t pad[3];
};

struct DecimFastArgs {
    DecimArgs d;
    uint32_t tb;          // threads of a CTA that own outputs (a multiple of 32, <= the CTA size): a CTA step makes 4 * tb outputs
    uint32_t lead;        // staged sample 0 is call-relative sample i0 - lead, so that it is a multiple of 4 (16-byte loads)
    uint32_t shape;       // min(3, (ntaps - 1) / M): selects the kernel instantiation
    uint32_t rm_magic;    // j / 4M == umulhi(j, rm_magic) for the staged indices of a step
    uint32_t tab_cap;     // phasor-table entries the launch reserves in shared memory (0: tables stay in global memory)
    uint32_t cuts[8];     // the walk's segment bounds: {0, M, 2M, 3M, ntaps, M + ntaps, 2M + ntaps, 3M + ntaps} sorted;
                          // segment i = positions [cuts[i], cuts[i+1])
    DfTaps tq[kDfMaxTq];
};

// 32-bit shared-window addresses: no generic-to-shared conversion inside the loops
__device__ __forceinline__ uint64_t lds_u64(uint32_t addr)
{
    uint64_t v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32x2(uint32_t addr, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
// the shared window's base as an opaque value (ptxas otherwise re-derives it from the cluster CTA id at every use: three issue
// slots per staged group)
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// The walk's seven segments.  Output k is active at positions [(3 - k) * M, (3 - k) * M + ntaps); the eight bounds sorted are
// the segments' cuts, and which outputs are active between two neighbouring cuts depends only on
// SHAPE = min(3, (ntaps - 1) / M) -- so the ranges are compile-time and the dispatch is straight-line code (a switch on a
// run-time range compiles to an indexed branch, after which ptxas no longer keeps the tap reads on the uniform datapath).
// Ties between cuts make empty segments.
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

// Walk positions [u, ue) with outputs KLO .. KHI active; pbase = shared address of the thread's base slot.  The padding slots
// are folded into a per-position offset (DfTaps::off, one more uniform load per position), so a segment is ONE loop whatever
// padding it crosses.  (A first form cut the segments into runs at the padding slots and stepped a pointer: the 1 .. 3-position
// remainders of those runs each exposed a full shared-memory latency -- 5-11 % slower, interleaved A/B.)  Software-pipelined:
// the next four samples are read before the current four are consumed (another 2-6 %).  u is a signed int so that the tap
// addresses of an unrolled body fold into immediate offsets of one uniform register.
template <int KLO, int KHI>
__device__ __forceinline__ void df_walk(const DecimFastArgs& A, uint32_t pbase, int u, int ue, uint64_t (&acc)[kDfR])
{
    auto load4 = [&](uint64_t (&y)[4], int at) {
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = lds_u64(pbase + A.tq[at + i].off);
    };
    auto fma4 = [&](const uint64_t (&y)[4], int at) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int k = KLO; k <= KHI; ++k) acc[k] = fma_f32x2(A.tq[at + i].h[k], y[i], acc[k]);
        }
    };
    if (u + 4 <= ue) {
        uint64_t ya[4], yb[4];
        load4(ya, u);
        while (u + 12 <= ue) {
            load4(yb, u + 4);
            fma4(ya, u);
            load4(ya, u + 8);
            fma4(yb, u + 4);
            u += 8;
        }
        if (u + 8 <= ue) {
            load4(yb, u + 4);
            fma4(ya, u);
            fma4(yb, u + 4);
            u += 8;
        } else {
            fma4(ya, u);
            u += 4;
        }
    }
    for (; u < ue; ++u) {
        const uint64_t y = lds_u64(pbase + A.tq[u].off);
#pragma unroll
        for (int k = KLO; k <= KHI; ++k) acc[k] = fma_f32x2(A.tq[u].h[k], y, acc[k]);
    }
}

template <int SHAPE>
__device__ __forceinline__ void df_walk_all(const DecimFastArgs& A, uint32_t pbase, uint64_t (&acc)[kDfR])
{
    df_walk<DfRange<SHAPE, 0>::lo, DfRange<SHAPE, 0>::hi>(A, pbase, (int)A.cuts[0], (int)A.cuts[1], acc);
    df_walk<DfRange<SHAPE, 1>::lo, DfRange<SHAPE, 1>::hi>(A, pbase, (int)A.cuts[1], (int)A.cuts[2], acc);
    df_walk<DfRange<SHAPE, 2>::lo, DfRange<SHAPE, 2>::hi>(A, pbase, (int)A.cuts[2], (int)A.cuts[3], acc);
    df_walk<DfRange<SHAPE, 3>::lo, DfRange<SHAPE, 3>::hi>(A, pbase, (int)A.cuts[3], (int)A.cuts[4], acc);
    df_walk<DfRange<SHAPE, 4>::lo, DfRange<SHAPE, 4>::hi>(A, pbase, (int)A.cuts[4], (int)A.cuts[5], acc);
    df_walk<DfRange<SHAPE, 5>::lo, DfRange<SHAPE, 5>::hi>(A, pbase, (int)A.cuts[5], (int)A.cuts[6], acc);
    df_walk<DfRange<SHAPE, 6>::lo, DfRange<SHAPE, 6>::hi>(A, pbase, (int)A.cuts[6], (int)A.cuts[7], acc);
}

// one 16-byte group (4 i16 or 2 f32 samples) against table entries ph, ph + 1, ...: mixed samples to shared address dst
template <int IN, bool TAB_SMEM>
__device__ __forceinline__ void df_mix_group(const uint4& raw, const float2* tab, uint32_t tab_addr, uint32_t ph, uint32_t dst)
{
    if constexpr (IN == C32) {   // already mixed: staged as they are
        sts_f32x2(dst, make_float2(__uint_as_float(raw.x), __uint_as_float(raw.y)));
        sts_f32x2(dst + 8u, make_float2(__uint_as_float(raw.z), __uint_as_float(raw.w)));
        return;
    }
    constexpr int V = IN == I16 ? 4 : 2;
    float2 phs[V];
#pragma unroll
    for (int s = 0; s < V; s++) {
        if constexpr (TAB_SMEM)
            phs[s] = lds_f32x2(tab_addr + (ph + s) * 8u);
        else
            phs[s] = __ldg(tab + ph + s);
    }
    if constexpr (IN == I16) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int s = 0; s < 4; s++) sts_f32x2(dst + 8u * s, cmul_unfused_fma(ingest_i16(w[s]), phs[s]));
    } else {
        sts_f32x2(dst, cmul_unfused_fma(make_float2(__uint_as_float(raw.x), __uint_as_float(raw.y)), phs[0]));
        sts_f32x2(dst + 8u, cmul_unfused_fma(make_float2(__uint_as_float(raw.z), __uint_as_float(raw.w)), phs[1]));
    }
}

// Phase A inside one tabled piece.  Group g of the step (16 bytes: 4 i16 or 2 f32 samples) belongs to thread g mod NT; a BATCH is
// UNR groups per thread.  The loads of batch b + 1 are issued before batch b is mixed (two register buffers), and batch 0 of
// the NEXT step before this step's walk (df_prefetch), so a CTA keeps global loads in flight through both of its phases.
constexpr int kDfUnr = 4;

template <int NT>
__device__ __forceinline__ void df_load_batch(uint4 (&r)[kDfUnr], const unsigned char* src, uint32_t batch, uint32_t ngroups, uint32_t tid)
{
    const unsigned char* s = src + (size_t)batch * (kDfUnr * NT * 16);
    const uint32_t g = batch * (kDfUnr * NT) + tid;
#pragma unroll
    for (int x = 0; x < kDfUnr; x++)
        if (g + x * NT < ngroups) r[x] = __ldcs(reinterpret_cast<const uint4*>(s + x * NT * 16));
}

template <int IN, bool TAB_SMEM, int NT>
__device__ __forceinline__ void df_stage_fast(const DecimFastArgs& A, const DevPiece& p, const unsigned char* src, int64_t ia, uint32_t ngroups,
                                              uint32_t ys_addr, uint32_t tab_addr, uint32_t tid, uint4 (&bufA)[kDfUnr], uint4 (&bufB)[kDfUnr],
                                              bool have_batch0)
{
    constexpr uint32_t V = IN == I16 ? 4 : 2;
    constexpr int UNR = kDfUnr;
    const float2* tab = IN == C32 ? nullptr : A.d.mix.tables + p.tab;
    const uint32_t period = IN == C32 ? 1u : p.period;
    uint32_t ph = IN == C32 ? 0u : piece_samplenum(p, (uint32_t)ia + tid * V - p.k_begin) - 1u;   // table phase of this thread's next group
    const uint32_t ph_step = IN == C32 ? 0u : (NT * V) % period;
    uint32_t j = tid * V;                                                         // staged index of this thread's next group
    auto mix_batch = [&](const uint4 (&r)[UNR], uint32_t batch) {
        const uint32_t g = batch * (UNR * NT) + tid;
#pragma unroll
        for (int x = 0; x < UNR; x++) {
            if (g + x * NT < ngroups) {
                df_mix_group<IN, TAB_SMEM>(r[x], tab, tab_addr, ph, ys_addr + (j + __umulhi(j, A.rm_magic)) * 8u);
                j += NT * V;
                if constexpr (IN != C32) {
                    ph += ph_step;
                    if (ph >= period) ph -= period;
                }
            }
        }
    };
    const uint32_t nb = (ngroups + UNR * NT - 1) / (UNR * NT);
    if (!have_batch0) df_load_batch<NT>(bufA, src, 0, ngroups, tid);
    for (uint32_t b = 0; b < nb; b += 2) {
        if (b + 1 < nb) df_load_batch<NT>(bufB, src, b + 1, ngroups, tid);
        mix_batch(bufA, b);
        if (b + 1 < nb) {
            if (b + 2 < nb) df_load_batch<NT>(bufA, src, b + 2, ngroups, tid);
            mix_batch(bufB, b + 1);
        }
    }
}

template <int IN, int OUT, int SHAPE, int NT>
__global__ void __launch_bounds__(NT) mix_decimate_fast_kernel(const __grid_constant__ DecimFastArgs A)
{
    constexpr int R = kDfR;
    constexpr uint32_t V = IN == I16 ? 4 : 2;             // samples per 16-byte load
    constexpr uint32_t kInBps = IN == I16 ? 4 : 8;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t tab_addr = opaque_u32(smem_u32(smem));   // [phasor table: tab_cap entries][staged samples]
    const uint32_t ys_addr = tab_addr + A.tab_cap * 8u;
    float2* tab_s = reinterpret_cast<float2*>(smem);
    float2* y_s = tab_s + A.tab_cap;
    const DecimArgs& d = A.d;
    const uint32_t M = d.M, RM = R * M, nh = d.ntaps - 1u, OT = A.tb * R, tid = threadIdx.x;
    const unsigned char* gin = static_cast<const unsigned char*>(d.mix.in);
    const uint32_t warp0 = __shfl_sync(0xffffffffu, tid & ~31u, 0);   // first thread of this warp, known uniform to the compiler
    uint32_t pi = 0, tab_piece = kNoPiece;                             // tab_piece: the piece whose table is in shared memory
    DevPiece p = get_piece(d.mix, 0);
    uint4 bufA[kDfUnr], bufB[kDfUnr];
    bool have_batch0 = false;   // bufA holds batch 0 of the coming step (loaded before the previous step's walk)
    // staged index j of a step holds call-relative sample ia + j (negative: history); ia is a multiple of 4
    auto step_origin = [&](uint32_t ob) { return (int64_t)d.first_out + (int64_t)ob * M - (int64_t)nh - (int64_t)A.lead; };
    auto step_groups = [&](uint32_t ob) {
        const uint32_t oc = d.nout - ob < OT ? d.nout - ob : OT;
        return (A.lead + (oc - 1u) * M + d.ntaps + V - 1u) / V;
    };
    auto in_input = [&](int64_t ia, uint32_t ngroups) { return ia >= 0 && (uint64_t)ia + (uint64_t)ngroups * V <= (uint64_t)d.mix.nsamples; };
    for (uint32_t ob = blockIdx.x * OT; ob < d.nout; ob += gridDim.x * OT) {
        const uint32_t ocount = d.nout - ob < OT ? d.nout - ob : OT;
        const int64_t ia = step_origin(ob);
        const uint32_t count = A.lead + (ocount - 1u) * M + d.ntaps;
        const uint32_t ngroups = step_groups(ob);
        // ---- phase A: mix the step's samples into shared memory
        bool fast = false;
        if constexpr (IN == C32) {
            fast = in_input(ia, ngroups);   // no pieces, no table: the mixer has been here already
        } else if (in_input(ia, ngroups)) {
            const uint32_t first = (uint32_t)ia, last = first + ngroups * V;
            if (first >= p.k_end || first < p.k_begin) {
                pi = find_piece(d.mix, first < p.k_begin ? 0u : pi, first);
                p = get_piece(d.mix, pi);
            }
            fast = last <= p.k_end && p.tab != kNoTab;
        }
        const bool tab_smem = fast && p.period + (uint32_t)kTabPad <= A.tab_cap;
        if (tab_smem && tab_piece != pi) {   // (block-uniform) first step inside this piece: copy its table
            __syncthreads();                 // the previous table's readers are done
            const float2* tab = d.mix.tables + p.tab;
            for (uint32_t e = tid; e < p.period + (uint32_t)kTabPad; e += NT) tab_s[e] = __ldg(tab + e);
            tab_piece = pi;
        }
        __syncthreads();   // the previous step's walks are done (and the table is staged)
        const unsigned char* src = gin + ((size_t)ia + (size_t)tid * V) * kInBps;   // (only dereferenced on the fast paths)
        if (tab_smem) {
            df_stage_fast<IN, true, NT>(A, p, src, ia, ngroups, ys_addr, tab_addr, tid, bufA, bufB, have_batch0);
        } else if (fast) {
            df_stage_fast<IN, false, NT>(A, p, src, ia, ngroups, ys_addr, tab_addr, tid, bufA, bufB, have_batch0);
        } else {
            for (uint32_t j = tid; j < count; j += NT) {
                const int64_t i = ia + (int64_t)j;
                float2 y = make_float2(0.0f, 0.0f);   // before the history / past the input: never read by a valid output
                if (i < 0) {
                    if (IN != C32 && i + (int64_t)nh >= 0) y = d.hist[(int64_t)nh + i];   // (C32: the history is part of the buffer)
                } else if (i < (int64_t)d.mix.nsamples) {
                    if constexpr (IN == C32)
                        y = __ldcs(reinterpret_cast<const float2*>(gin) + i);
                    else
                        y = mix_one<IN>(d.mix, (uint32_t)i, pi, p);
                }
                y_s[j + j / RM] = y;
            }
        }
        // batch 0 of this CTA's next step goes out now: its latency hides behind the walk
        have_batch0 = false;
        {
            const uint64_t ob_next = (uint64_t)ob + (uint64_t)gridDim.x * OT;
            if (ob_next < d.nout) {
                const int64_t ia_n = step_origin((uint32_t)ob_next);
                const uint32_t ng_n = step_groups((uint32_t)ob_next);
                if (in_input(ia_n, ng_n)) {
                    df_load_batch<NT>(bufA, gin + ((size_t)ia_n + (size_t)tid * V) * kInBps, 0, ng_n, tid);
                    have_batch0 = true;
                }
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
            df_walk_all<SHAPE>(A, ys_addr + tid * (RM + 1u) * 8u, acc);
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

using DecimFastKernel = void (*)(const DecimFastArgs);
// one translation unit per type pair (decimate_inst.cu)
DecimFastKernel df_kernel_0_0(int nt128, int shape);
DecimFastKernel df_kernel_0_1(int nt128, int shape);
DecimFastKernel df_kernel_1_0(int nt128, int shape);
DecimFastKernel df_kernel_1_1(int nt128, int shape);
DecimFastKernel df_kernel_2_0(int nt128, int shape);   // already-mixed input (two-pass form), i16 / f32 output
DecimFastKernel df_kernel_2_1(int nt128, int shape);

}  // namespace dmix
