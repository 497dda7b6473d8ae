// tools/tune/tune.cu -- standalone tuning harness for the mixer's memory structure (NOT product code).
//
// Times, with CUDA events on the launch stream, (a) the product kernel template
// dmix::mix_kernel<IN, OUT, T, U, MINB> under different CTA shapes / schedules, (b) a cast-only
// pass-through of the same loop (what the load/store structure can reach with no phasor math),
// (c) a 1-D bulk-async (TMA, cp.async.bulk + mbarrier) copy pipeline as the ceiling for a
// shared-memory-staged design.  Prints one JSON line per variant.
//
// Build: make -C tools/tune        Run (GPU box): tools/tune/tune [filter-substring] [big]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../doppler_b200/csrc/mixer_kernels.cuh"

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)


// ---------------------------------------------------------------------------------------------
// The round-1 "v0" register-staged kernel (LDG -> registers -> STG), kept here only as the baseline
// the bulk-async pipelines are measured against; it is not part of the product any more.
namespace dmix {
constexpr int kUnroll = 4;
__host__ __device__ constexpr uint32_t tile_samples(int in, int out, int threads = kThreads, int unroll = kUnroll)
{
    return (uint32_t)threads * unroll * group_samples(in, out);
}
enum LegacyMode { kLegacyTabShared = 0, kLegacyTabGlobal = 1, kDirectPeriodic = 2, kDirectLinear = 3 };
struct LegacyArgs : MixArgs {
    uint32_t ntiles, tiles_per_cta, smem_entries, interleave;
};
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

// A full tile inside one piece.  All U group loads are issued before any arithmetic.
template <int IN, int OUT, int MODE, int T, int U>
__device__ __forceinline__ void fast_tile(const LegacyArgs& a, const DevPiece& p, uint32_t k0, const float2* tab)
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
            if constexpr (MODE == kLegacyTabShared) {
                ph = tab[j + s];                    // padded: no wrap inside a group
            } else if constexpr (MODE == kLegacyTabGlobal) {
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
__device__ __noinline__ void slow_tile(const LegacyArgs& a, uint32_t pi, uint32_t k0)
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
__global__ void __launch_bounds__(T, MINB) mix_kernel(const __grid_constant__ LegacyArgs a)
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
            fast_tile<IN, OUT, kLegacyTabShared, T, U>(a, p, k0, tab_s);
        } else {
            fast_tile<IN, OUT, kLegacyTabGlobal, T, U>(a, p, k0, a.tables + p.tab);
        }
    }
}


}  // namespace dmix

using dmix::DevPiece;
using dmix::MixArgs;

static double g_peak = 6550.1;
static int g_sms = 148;
static const char* g_filter = nullptr;
static cudaStream_t g_stream;

static void magic_for(uint32_t d, uint32_t* magic, uint32_t* shift)
{
    uint32_t l = 0;
    while ((1ull << l) < d) l++;
    *shift = 31 + l;
    *magic = (uint32_t)(((1ull << (31 + l)) / d) + 1);
}

// ---------------------------------------------------------------------------------------------
// (b) cast-only pass-through with the mixer's tile loop
template <int IN, int OUT, int T, int U>
__global__ void __launch_bounds__(T) passthru_kernel(const void* in, void* out, uint32_t ntiles)
{
    constexpr int G = dmix::group_samples(IN, OUT);
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t g0 = tile * (T * U) + threadIdx.x;
        float2 smp[U][G];
#pragma unroll
        for (int u = 0; u < U; u++) dmix::load_group<IN, G>(in, g0 + u * T, smp[u]);
#pragma unroll
        for (int u = 0; u < U; u++) dmix::store_group<OUT, G>(out, g0 + u * T, smp[u]);
    }
}

// ---------------------------------------------------------------------------------------------
// (c) bulk-async copy pipeline: one thread per CTA drives a ring of STAGES tiles
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

// Every warp of the CTA runs its own independent pipeline (lane 0 issues) over its own ring of
// STAGES buffers; pipelines are numbered blockIdx.x * nwarps + warp.  contig = 1: pipeline p owns
// tiles [p*per, (p+1)*per); contig = 0: tiles p, p + npipes, ...
template <int STAGES>
__global__ void __launch_bounds__(1024) bulk_copy_kernel(const char* in, char* out, uint32_t tile_bytes, uint32_t ntiles, int contig)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full_all[32 * STAGES];
    if ((threadIdx.x & 31) != 0) return;
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint64_t* full = full_all + warp * STAGES;
    char* ring = smem + (size_t)warp * STAGES * tile_bytes;
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t pipe = blockIdx.x * nwarps + warp, npipes = gridDim.x * nwarps;
    uint32_t first, step, mine;
    if (contig) {
        const uint32_t per = (ntiles + npipes - 1) / npipes;
        first = pipe * per;
        step = 1;
        mine = first < ntiles ? min(per, ntiles - first) : 0;
    } else {
        first = pipe;
        step = npipes;
        mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    }
    auto load = [&](uint32_t i) {
        const uint32_t s = i % STAGES;
        mbar_expect_tx(&full[s], tile_bytes);
        bulk_g2s(ring + (size_t)s * tile_bytes, in + (size_t)(first + i * step) * tile_bytes, tile_bytes, &full[s]);
    };
    for (uint32_t i = 0; i < (uint32_t)(STAGES - 1) && i < mine; i++) load(i);
    for (uint32_t i = 0; i < mine; i++) {
        const uint32_t s = i % STAGES;
        mbar_wait(&full[s], (i / STAGES) & 1u);
        bulk_s2g(out + (size_t)(first + i * step) * tile_bytes, ring + (size_t)s * tile_bytes, tile_bytes);
        bulk_commit();
        if (i + STAGES - 1 < mine) {
            bulk_wait_read<1>();   // the store issued one iteration ago has drained its buffer
            load(i + STAGES - 1);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
struct Timing {
    double med_us, best_us;
};

template <typename F>
static Timing time_it(F&& launch, int iters = 15)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<double> t;
    if (getenv("TUNE_SUSTAINED")) {
        // bench.py's regime: launches queued back to back under load (the 1 kW power cap pulls the SM clock to
        // ~1.65 GHz after a few hundred ms of streaming) -- 0.4 s of warm-up launches, then batches of 10
        auto t0 = std::chrono::steady_clock::now();
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 0.4) {
            for (int i = 0; i < 20; i++) launch();
            CK(cudaStreamSynchronize(g_stream));
        }
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaEventRecord(e0, g_stream));
            for (int i = 0; i < 10; i++) launch();
            CK(cudaEventRecord(e1, g_stream));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            t.push_back(ms * 1e3 / 10);
        }
        std::sort(t.begin(), t.end());
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return Timing{t[t.size() / 2], t[0]};
    }
    for (int i = 0; i < 3 + iters; i++) {
        CK(cudaEventRecord(e0, g_stream));
        launch();
        CK(cudaEventRecord(e1, g_stream));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i >= 3) t.push_back(ms * 1e3);
    }
    std::sort(t.begin(), t.end());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return Timing{t[t.size() / 2], t[0]};
}

static void report(const std::string& name, uint64_t nsamples, double bytes_per_sample, Timing tm, const std::string& extra = "")
{
    const double gbs = nsamples * bytes_per_sample / (tm.med_us * 1e-6) / 1e9;
    printf("{\"variant\": \"%s\", \"samples\": %llu, \"median_us\": %.1f, \"best_us\": %.1f, \"gbs\": %.1f, \"frac\": %.4f%s}\n", name.c_str(),
           (unsigned long long)nsamples, tm.med_us, tm.best_us, gbs, gbs / g_peak, extra.c_str());
    fflush(stdout);
}

static bool want(const std::string& name) { return !g_filter || strstr(name.c_str(), g_filter) != nullptr; }

static const char* tname(int t) { return t == 0 ? "i16" : "f32"; }

struct Buffers {
    void *in = nullptr, *out = nullptr;
    float2* tab = nullptr;
    uint64_t cap_in = 0, cap_out = 0;
};
static Buffers g_buf;

__global__ void fill_kernel(uint32_t* p, uint64_t nwords, int f32)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nwords; i += stride) {
        uint32_t h = (uint32_t)i * 2654435761u ^ (uint32_t)(i >> 32) * 40503u;
        h ^= h >> 15;
        h *= 2246822519u;
        h ^= h >> 13;
        if (f32) {
            const float v = ((int)(h & 0xffffff) - 0x800000) * (0.7f / 8388608.0f);
            p[i] = __float_as_uint(v);
        } else {
            const int a = (int)(h & 0x7fff) - 16384, b = (int)((h >> 16) & 0x7fff) - 16384;
            p[i] = (uint32_t)(uint16_t)(short)a | ((uint32_t)(uint16_t)(short)b << 16);
        }
    }
}

template <int IN, int OUT, int T, int U, int MINB>
static void run_mix(uint64_t n, int interleave, int ctas_per_sm /*0 = occupancy*/, uint32_t period, float r)
{
    char name[160];
    snprintf(name, sizeof name, "mix %s->%s T%d U%d minb%d %s cps%d P%u", tname(IN), tname(OUT), T, U, MINB, interleave ? "ilv" : "contig",
             ctas_per_sm, period);
    if (!want(name)) return;
    auto kern = dmix::mix_kernel<IN, OUT, T, U, MINB>;
    const size_t smem = (period + dmix::kTabPad) * sizeof(float2);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    constexpr int G = dmix::group_samples(IN, OUT);
    const uint32_t tile = dmix::tile_samples(IN, OUT, T, U);
    dmix::LegacyArgs a;
    memset(&a, 0, sizeof a);
    a.in = g_buf.in;
    a.out = g_buf.out;
    a.tables = g_buf.tab;
    a.nsamples = (uint32_t)n;
    a.npieces = 1;
    a.ntiles = (uint32_t)((n + tile - 1) / tile);
    a.smem_entries = period;
    a.interleave = interleave;
    DevPiece d;
    memset(&d, 0, sizeof d);
    d.k_begin = 0;
    d.k_end = (uint32_t)n;
    d.base = 0;
    d.period = period;
    d.r = r;
    d.tab = 0;
    magic_for(period, &d.magic, &d.shift);
    d.step_u = (uint32_t)((uint64_t)(T * G) % period);
    a.inl[0] = d;
    const int cps = ctas_per_sm ? ctas_per_sm : occ;
    uint32_t grid;
    if (interleave) {
        grid = std::min<uint32_t>(a.ntiles, (uint32_t)(g_sms * cps));
        a.tiles_per_cta = 0;
    } else {
        const uint32_t target = (uint32_t)(g_sms * cps * 4);
        a.tiles_per_cta = std::max<uint32_t>(1, (a.ntiles + target - 1) / target);
        grid = (a.ntiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
    }
    Timing tm = time_it([&] { kern<<<grid, T, smem, g_stream>>>(a); });
    char extra[128];
    snprintf(extra, sizeof extra, ", \"regs\": %d, \"occ\": %d, \"grid\": %u", fa.numRegs, occ, grid);
    report(name, n, (IN ? 8 : 4) + (OUT ? 8 : 4), tm, extra);
}

template <int IN, int OUT, int T, int U>
static void run_passthru(uint64_t n, int ctas_per_sm)
{
    char name[160];
    snprintf(name, sizeof name, "passthru %s->%s T%d U%d cps%d", tname(IN), tname(OUT), T, U, ctas_per_sm);
    if (!want(name)) return;
    auto kern = passthru_kernel<IN, OUT, T, U>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, 0));
    const uint32_t tile = dmix::tile_samples(IN, OUT, T, U);
    const uint32_t ntiles = (uint32_t)(n / tile);
    const int cps = ctas_per_sm ? ctas_per_sm : occ;
    const uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)(g_sms * cps));
    Timing tm = time_it([&] { kern<<<grid, T, 0, g_stream>>>(g_buf.in, g_buf.out, ntiles); });
    char extra[64];
    snprintf(extra, sizeof extra, ", \"occ\": %d, \"grid\": %u", occ, grid);
    report(name, (uint64_t)ntiles * tile, (IN ? 8 : 4) + (OUT ? 8 : 4), tm, extra);
}

template <int STAGES>
static void run_bulk(uint64_t bytes, uint32_t tile_bytes, int ctas_per_sm, int nwarps = 1, int contig = 0)
{
    char name[160];
    snprintf(name, sizeof name, "bulkcopy tile%uK stages%d cps%d warps%d %s bytes%lluM", tile_bytes >> 10, STAGES, ctas_per_sm, nwarps,
             contig ? "contig" : "ilv", (unsigned long long)(bytes >> 20));
    if (!want(name)) return;
    auto kern = bulk_copy_kernel<STAGES>;
    const size_t smem = (size_t)STAGES * tile_bytes * nwarps;
    if (smem > 200 * 1024) return;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * nwarps, smem));
    if (occ < ctas_per_sm) {
        printf("{\"variant\": \"%s\", \"skipped\": \"occupancy %d < %d\"}\n", name, occ, ctas_per_sm);
        return;
    }
    const uint32_t ntiles = (uint32_t)(bytes / tile_bytes);
    const uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)(g_sms * ctas_per_sm));
    Timing tm = time_it([&] { kern<<<grid, 32 * nwarps, smem, g_stream>>>((const char*)g_buf.in, (char*)g_buf.out, tile_bytes, ntiles, contig); });
    char extra[64];
    snprintf(extra, sizeof extra, ", \"inflight_kb_per_sm\": %d", (int)((STAGES - 1) * (tile_bytes >> 10) * nwarps * ctas_per_sm));
    // report as "samples" of 8 B (i16->i16 equivalent): bytes moved = 2 * bytes
    report(name, (uint64_t)ntiles * tile_bytes / 4, 8, tm, extra);
}

template <int IN, int OUT, int WARPS, int S, int U>
static void run_stream(uint64_t n, uint32_t period, float r, int tabmode /*0 smem, 1 global(L2), 2 direct*/)
{
    using C = dmix::StreamCfg<IN, OUT, WARPS, S, U>;
    char name[160];
    snprintf(name, sizeof name, "stream %s->%s W%d S%d U%d %s P%u", tname(IN), tname(OUT), WARPS, S, U,
             tabmode == 0 ? "smemtab" : tabmode == 1 ? "l2tab" : "direct", period);
    if (!want(name)) return;
    auto kern = dmix::mix_grid_kernel<IN, OUT, WARPS, S, U>;
    const size_t smem = C::kGridSmem + (tabmode == 0 ? C::table_bytes(period) : 0);
    if (smem > 227 * 1024) return;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    MixArgs a;
    memset(&a, 0, sizeof a);
    a.in = g_buf.in;
    a.out = g_buf.out;
    a.tables = g_buf.tab;
    a.nsamples = (uint32_t)n;
    a.npieces = 1;
    const uint32_t ntiles = (uint32_t)(n / C::kTileSamples);
    a.nsegs = 1;
    a.nunits = ntiles;
    a.tail_begin = ntiles * C::kTileSamples;   // one GRID segment of whole tiles
    a.inl_segs[0].unit_begin = 0;
    a.inl_segs[0].unit_end = ntiles;
    a.inl_segs[0].k_begin = 0;
    a.inl_segs[0].k_end = a.tail_begin;
    a.smem_piece = tabmode == 0 ? 0 : dmix::kNoPiece;
    DevPiece d;
    memset(&d, 0, sizeof d);
    d.k_begin = 0;
    d.k_end = (uint32_t)n;
    d.base = 0;
    d.period = period;
    d.r = r;
    d.tab = tabmode == 2 ? dmix::kNoTab : 0;
    if (period) {
        magic_for(period, &d.magic, &d.shift);
        d.step_u = (uint32_t)C::kRow % period;
    }
    a.inl[0] = d;
    const uint32_t grid = std::min<uint32_t>((uint32_t)g_sms, (ntiles + WARPS - 1) / WARPS);
    Timing tm = time_it([&] { kern<<<grid, WARPS * 32, smem, g_stream>>>(a); });
    char extra[160];
    snprintf(extra, sizeof extra, ", \"regs\": %d, \"smem\": %d, \"inflight_kb_per_sm\": %d", fa.numRegs, (int)smem, WARPS * S * C::kTileIn / 1024);
    report(name, n, (IN ? 8 : 4) + (OUT ? 8 : 4), tm, extra);
}

template <int IN, int OUT, int WARPS>
static void sweep_stream_w(uint64_t n)
{
    const uint32_t P = 256;
    const float r = -15000.0f / 256000.0f;
    run_stream<IN, OUT, WARPS, 2, 1>(n, P, r, 0);
    run_stream<IN, OUT, WARPS, 3, 1>(n, P, r, 0);
    run_stream<IN, OUT, WARPS, 4, 1>(n, P, r, 0);
    run_stream<IN, OUT, WARPS, 2, 2>(n, P, r, 0);
    run_stream<IN, OUT, WARPS, 3, 2>(n, P, r, 0);
    run_stream<IN, OUT, WARPS, 4, 2>(n, P, r, 0);
    run_stream<IN, OUT, WARPS, 2, 3>(n, P, r, 0);
    run_stream<IN, OUT, WARPS, 3, 3>(n, P, r, 0);
}

template <int IN, int OUT>
static void sweep_stream(uint64_t n)
{
    sweep_stream_w<IN, OUT, 12>(n);
    sweep_stream_w<IN, OUT, 16>(n);
    sweep_stream_w<IN, OUT, 20>(n);
    sweep_stream_w<IN, OUT, 24>(n);
    sweep_stream_w<IN, OUT, 28>(n);
    sweep_stream_w<IN, OUT, 32>(n);
}

// direct (table-free) evaluation on a piece that never resets, r = 1 Hz / 2 Gsps: per-tile pipeline overhead
// against tile size, lean loop
template <int IN, int OUT>
static void sweep_direct(uint64_t n)
{
    const float r = 1.0f / 2.0e9f;
    run_stream<IN, OUT, 20, 2, 2>(n, 0, r, 2);
    run_stream<IN, OUT, 16, 2, 3>(n, 0, r, 2);
    run_stream<IN, OUT, 16, 2, 4>(n, 0, r, 2);
    run_stream<IN, OUT, 12, 2, 4>(n, 0, r, 2);
    run_stream<IN, OUT, 12, 2, 6>(n, 0, r, 2);
    run_stream<IN, OUT, 16, 2, 6>(n, 0, r, 2);
    run_stream<IN, OUT, 8, 2, 8>(n, 0, r, 2);
    run_stream<IN, OUT, 12, 2, 8>(n, 0, r, 2);
}

template <int IN, int OUT>
static void sweep_pair(uint64_t n)
{
    const uint32_t P = 256;
    const float r = -15000.0f / 256000.0f;
    // current product shape, then schedule / shape variants
    run_mix<IN, OUT, 256, 4, 0>(n, 0, 0, P, r);
    run_mix<IN, OUT, 256, 4, 0>(n, 1, 0, P, r);
    run_mix<IN, OUT, 256, 2, 0>(n, 1, 0, P, r);
    run_mix<IN, OUT, 256, 8, 0>(n, 1, 0, P, r);
    run_mix<IN, OUT, 512, 2, 0>(n, 1, 0, P, r);
    run_mix<IN, OUT, 512, 4, 0>(n, 1, 0, P, r);
    run_mix<IN, OUT, 128, 4, 0>(n, 1, 0, P, r);
    run_mix<IN, OUT, 128, 8, 0>(n, 1, 0, P, r);
    run_mix<IN, OUT, 256, 4, 5>(n, 1, 0, P, r);
    run_mix<IN, OUT, 256, 4, 6>(n, 1, 0, P, r);
    run_mix<IN, OUT, 256, 2, 8>(n, 1, 0, P, r);
    run_mix<IN, OUT, 256, 4, 0>(n, 1, 2, P, r);
    run_mix<IN, OUT, 256, 4, 0>(n, 1, 3, P, r);
    run_mix<IN, OUT, 256, 8, 0>(n, 1, 2, P, r);
    run_passthru<IN, OUT, 256, 4>(n, 0);
    run_passthru<IN, OUT, 256, 8>(n, 0);
    run_passthru<IN, OUT, 512, 4>(n, 0);
    run_passthru<IN, OUT, 256, 4>(n, 4);
    run_passthru<IN, OUT, 256, 2>(n, 0);
}

int main(int argc, char** argv)
{
    if (argc > 1 && strcmp(argv[1], "all") != 0) g_filter = argv[1];
    const bool big = argc > 2 && strcmp(argv[2], "big") == 0;
    if (const char* p = getenv("DOPPLER_PEAK_GBS")) g_peak = atof(p);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    fprintf(stderr, "device %s, %d SMs, peak %.1f GB/s\n", prop.name, g_sms, g_peak);

    const uint64_t n_small = getenv("TUNE_N") ? strtoull(getenv("TUNE_N"), nullptr, 10) : (256ull << 20);   // 268 M samples
    const uint64_t n_big = big ? (1000ull << 20) : std::max<uint64_t>(n_small, 256ull << 20);   // ~1.05 G samples (< 2^30)
    const uint64_t cap = n_big * 8;
    CK(cudaMalloc(&g_buf.in, cap));
    CK(cudaMalloc(&g_buf.out, cap));
    CK(cudaMalloc(&g_buf.tab, (4096 + 8) * sizeof(float2)));
    dmix::build_phasor_table_kernel<<<2, 256, 0, g_stream>>>(g_buf.tab, -15000.0f / 256000.0f, 256, 256 + dmix::kTabPad);
    CK(cudaStreamSynchronize(g_stream));

    for (int f32 = 0; f32 < 2; f32++) {
        fill_kernel<<<g_sms * 8, 256, 0, g_stream>>>((uint32_t*)g_buf.in, cap / 4, f32);
        CK(cudaStreamSynchronize(g_stream));
        if (f32 == 0) {
            sweep_pair<0, 0>(n_small);
            sweep_pair<0, 1>(n_small);
            sweep_stream<0, 0>(n_small);
            sweep_stream<0, 1>(n_small);
            sweep_direct<0, 0>(n_small);
            sweep_direct<0, 1>(n_small);
            if (big) {
                run_mix<0, 0, 256, 4, 0>(n_big, 0, 0, 256, -15000.0f / 256000.0f);
                run_mix<0, 0, 256, 4, 0>(n_big, 1, 0, 256, -15000.0f / 256000.0f);
                run_passthru<0, 0, 256, 4>(n_big, 0);
            }
        } else {
            sweep_pair<1, 0>(n_small);
            sweep_pair<1, 1>(n_small);
            sweep_stream<1, 0>(n_small);
            sweep_stream<1, 1>(n_small);
            sweep_direct<1, 0>(n_small);
            sweep_direct<1, 1>(n_small);
            if (big) {
                run_mix<1, 0, 256, 4, 0>(n_big, 0, 0, 256, -15000.0f / 256000.0f);
                run_mix<1, 0, 256, 4, 0>(n_big, 1, 0, 256, -15000.0f / 256000.0f);
            }
        }
    }
    // bulk-async copy ceilings: 1 GiB and (big) 4 GiB each way
    for (uint64_t bytes : {1ull << 30, big ? (4000ull << 20) : 0ull}) {
        if (!bytes) continue;
        for (uint32_t tk : {2u, 4u, 8u, 16u, 32u, 64u}) {
            run_bulk<2>(bytes, tk << 10, 1);
            run_bulk<3>(bytes, tk << 10, 1);
            run_bulk<4>(bytes, tk << 10, 1);
            run_bulk<6>(bytes, tk << 10, 1);
            run_bulk<8>(bytes, tk << 10, 1);
        }
        for (int contig : {0, 1}) {
            run_bulk<4>(bytes, 16 << 10, 1, 1, contig);
            run_bulk<2>(bytes, 16 << 10, 2, 1, contig);
            run_bulk<3>(bytes, 16 << 10, 2, 1, contig);
            run_bulk<2>(bytes, 8 << 10, 2, 1, contig);
            run_bulk<3>(bytes, 8 << 10, 2, 1, contig);
            run_bulk<2>(bytes, 8 << 10, 4, 1, contig);
            for (int nw : {2, 4, 8, 16}) {
                run_bulk<2>(bytes, 2 << 10, 1, nw, contig);
                run_bulk<3>(bytes, 2 << 10, 1, nw, contig);
                run_bulk<4>(bytes, 2 << 10, 1, nw, contig);
                run_bulk<2>(bytes, 4 << 10, 1, nw, contig);
                run_bulk<3>(bytes, 4 << 10, 1, nw, contig);
                run_bulk<4>(bytes, 4 << 10, 1, nw, contig);
                run_bulk<2>(bytes, 8 << 10, 1, nw, contig);
                run_bulk<3>(bytes, 8 << 10, 1, nw, contig);
            }
        }
    }
    return 0;
}
