// doppler_b200.cu -- C ABI (include/doppler_b200.h) over the sm_100a mixer kernels.
//
// Host side of the drop-in boundary: validates like the reference asserts (dsp.rs:87,103),
// plans the samplenum state machine analytically (plan.h), cuts every launch into GRID / COLUMN
// segments (mixer_kernels.cuh), keeps the phasor tables of short periods in a small device arena,
// and drives the kernels either on caller-owned device buffers or on host buffers through a
// 3-slot pinned/stream pipeline (H2D, kernel and D2H of neighbouring chunks overlap).  No CPU
// implementation of the mixer exists in this library.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/doppler_b200.h"
#include "mixer_kernels.cuh"
#include "decimate_kernels.cuh"
#include "plan.h"
#include "collect.h"

// (WARPS, S, U) of the segmented kernel per type pair, chosen on B200 (profiles/r01_seg_tune.md: warps
// a multiple of the 4 schedulers, 12-16 samples per lane per tile so that the per-tile pipeline overhead is
// amortised; i16->i16 is issue-bound and wants 16 warps, which leaves room for a ~2600-entry table only).  SEGV selects a tuning variant at build time
// (`make variants`, tools/gpu/gpu_seg_tune.sh); the product is SEGV 0.
#ifndef SEGV
#define SEGV 0
#endif
#if SEGV == 1
#define SEG_I16I16 16, 2, 4
#define SEG_I16F32 16, 3, 4
#define SEG_F32I16 16, 2, 4
#define SEG_F32F32 16, 2, 4
#elif SEGV == 2
#define SEG_I16I16 16, 2, 3
#define SEG_I16F32 16, 3, 3
#define SEG_F32I16 20, 2, 3
#define SEG_F32F32 16, 2, 3
#elif SEGV == 3
#define SEG_I16I16 16, 3, 3
#define SEG_I16F32 12, 3, 4
#define SEG_F32I16 16, 2, 4
#define SEG_F32F32 12, 2, 4
#else
#define SEG_I16I16 16, 2, 4
#define SEG_I16F32 12, 3, 4
#define SEG_F32I16 16, 2, 4
#define SEG_F32F32 12, 2, 4
#endif

using dmix::DevPiece;
using dmix::DevSeg;
using dmix::MixArgs;

namespace {

constexpr uint64_t kLaunchMaxSamples = 1ull << 30;   // k fits 32 bits with room for base + offset
constexpr uint32_t kColumnMaxRows = 64;              // COLUMN segments: rows sharing one phasor evaluation, at most
constexpr uint32_t kUnitsPerPipe = 4;                // ... halved once if the launch has fewer work units per pipeline than this
constexpr double kWindowCostTiles = 1.3;             // issue cost of one COLUMN window evaluation, in tiles (ncu: ~650 vs ~510 instructions)
constexpr uint32_t kColumnMinRows = 2;               // fewer whole periods than this: evaluate per sample instead
constexpr uint32_t kColumnMinLaunch = 4u << 20;      // launches below this many samples are latency-bound: a serial window
                                                     // evaluation per work unit costs more than it saves (measured 23 vs 13 us at 1 M)
constexpr size_t kArenaEntries = 4ull << 20;         // 32 MiB of (cos, sin) pairs: >= 1000 tables of the longest tabled period
constexpr uint32_t kSmemTabMaxEntries = 4096;        // 32 KiB of shared memory per CTA at most
constexpr size_t kHostChunkBytes = 32ull << 20;      // host-path pipeline chunk (input side)
constexpr int kSlots = 3;
constexpr uint32_t kSmallMaxSamples = 4u << 20;      // launches up to this size take the latency-shaped kernel (mix_small_kernel); tools/tune
constexpr size_t kTinyHostBytes = 128u << 10;
constexpr size_t kTinyStageBytes = 32u << 10;        // ... and up to this much is staged by memcpy without querying the caller's pointers        // host-buffer calls up to this much input run zero-copy over mapped pinned memory
constexpr int kMetaSlots = 8;                        // pinned staging slots for launch metadata

thread_local std::string g_create_error;   // create() errors, read back by the calling thread through last_error(NULL)

struct TableRef {
    uint32_t off;
    uint32_t period;
};

struct MetaSlot {
    cudaEvent_t copied = nullptr;   // the image has reached the device (recorded on the context's copy stream)
    void* host = nullptr;   // pinned staging image
    void* dev = nullptr;    // its device copy (persistent: stream-ordered pool memory is trimmed at every synchronize)
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool used = false;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* h_in = nullptr;    // pinned staging
    void* h_out = nullptr;
    void* h_in_dev = nullptr;   // device aliases of the staging buffers (zero-copy tiny path)
    void* h_out_dev = nullptr;
    size_t in_cap = 0, out_cap = 0;
    // pending copy-out of a staged result
    void* user_out = nullptr;
    size_t user_out_bytes = 0;
    bool busy = false;
};

}  // namespace

struct doppler_b200_ctx {
    int device = 0;
    void* copy_pool = nullptr;   // CopyPool (host staging threads), created on first use
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    dplan::Planner planner;
    float2* arena = nullptr;
    size_t arena_cap = 0, arena_used = 0;
    std::unordered_map<uint32_t, TableRef> tables;   // key: bits of r
    cudaEvent_t tables_ready = nullptr;
    bool tables_event_valid = false;
    cudaStream_t tables_stream = nullptr;   // the stream the latest table build (and tables_ready) was enqueued on
    Slot slots[kSlots];
    MetaSlot meta[kMetaSlots];   // pinned staging for launch metadata (pieces, segments, work-unit counter)
    uint32_t meta_next = 0;
    cudaStream_t meta_stream = nullptr;   // metadata uploads run here, beside the previous launch's kernel
    std::string err;
    uint64_t launches = 0;
    std::vector<const void*> configured;   // kernels whose dynamic shared-memory limit has been raised
    uint32_t small_max = kSmallMaxSamples;
    bool seg_alt = false;               // doppler_b200_tune: use StreamShape::seg_alt
    bool decim_generic = false;         // doppler_b200_tune: the fused decimator always takes the generic kernel
    uint32_t decim_stage_slots = dmix::kDfStageSlots;   // doppler_b200_tune: shared-memory slots of one CTA step of the register-blocked decimator
    uint32_t max_claim = 1;             // work units claimed at once by the segmented kernels (doppler_b200_tune; chunks of up to 8
                                        // lost the A/B by 6-15 %, profiles/r02_ab_seg.jsonl)
    size_t tiny_host_bytes = kTinyHostBytes;
    // zero-copy per-block host path: completion flag in mapped host memory + CTA counter on the device
    uint32_t* done_flag = nullptr;      // host (pinned)
    uint32_t* done_flag_dev = nullptr;  // its device alias
    uint32_t* done_counter = nullptr;   // device
    uint32_t done_token = 0;
    // resident kernel of the per-block host path (mixer_kernels.cuh: mix_resident_kernel)
    dmix::RtMailbox* rt_mb = nullptr;       // mapped pinned host memory
    dmix::RtMailbox* rt_mb_dev = nullptr;   // its device alias
    cudaStream_t rt_stream = nullptr;
    uint32_t* rt_out = nullptr;             // the result in flagged units {word, request number}: mapped pinned host memory
    void* rt_out_dev = nullptr;
    const void* rt_in_dev = nullptr;        // what the running kernel was launched with (a request with other addresses restarts it)
    const void* rt_tables_dev = nullptr;
    uint32_t rt_seq = 0, rt_gen = 0;        // rt_seq: the latest request, served in full once rt_request has returned
    uint64_t rt_idle_us = 20000;            // the kernel leaves after this long without a request; 0: no resident kernel (doppler_b200_tune)
    uint64_t rt_requests = 0, rt_starts = 0;
    // The steady state of a block stream with one shift -- samplenum inside its period, ONE periodic piece per block --
    // planned without the planner: what the last such block's plan looked like (tiny_host_call).
    struct RtSteady {
        bool valid = false, off = false;   // off: DOPPLER_B200_NO_STEADY_RULE=1 (A/B)
        uint32_t r_bits = 0;
        DevPiece piece;   // period, r, tab, magic, shift of that plan (k_begin .. base are per block)
    } rt_steady;
    // The ratio of the previous per-block host call (one shift): a block stream that keeps its ratio gets a phasor table even
    // when ONE block is shorter than two periods -- the table pays across the calls (launch_mix, tiny_host_call).
    bool tiny_last_r_valid = false;
    uint32_t tiny_last_r_bits = 0;
    uint64_t rt_traced = 0, rt_steady_traced = 0;
    uint64_t rt_steady_ns[3] = {0, 0, 0};   // the same for calls planned by the steady-state rule (tiny_host_call)
    uint64_t rt_ns[3] = {0, 0, 0};          // DOPPLER_B200_TRACE=1: stage in, plan, request -> result collected in the caller's buffer (resident-kernel calls only)
    // DOPPLER_B200_TRACE=1: phase clock of the tiny host path (ns totals: staging in, plan + launch, wait for the flag, copy out)
    bool trace = false;
    uint64_t tiny_calls = 0, tiny_ns[4] = {0, 0, 0, 0};
};

namespace {

int fail(doppler_b200_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail((ctx), DOPPLER_B200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                         \
    } while (0)

inline size_t bytes_per_sample(int t) { return t == DOPPLER_B200_I16 ? 4 : 8; }
inline bool valid_type(int t) { return t == DOPPLER_B200_I16 || t == DOPPLER_B200_F32; }

using MixKernel = void (*)(const MixArgs);

// One kernel's shape: WARPS pipelines per CTA, S stages, U rows per tile (tools/tune, profiles/).
struct KernShape {
    MixKernel kern;
    int warps;
    uint32_t tile_samples, row_samples, gran;
    uint32_t fixed_smem;
    uint32_t (*table_bytes)(uint32_t period);
    uint32_t smem_tab_entries;   // longest period whose table this kernel stages in shared memory
    uint32_t scratch_smem;       // per-CTA plateau scratch (direct-evaluation shape of the lean kernel), part of fixed_smem
    int pipes;                   // pipelines per CTA (= warps, or warps / Q for the Q-warp pipelines)
};
// Per (intype, outtype): the lean loop for a launch that is one GRID segment (const mode), and the
// segmented loop (GRID + COLUMN segments) with its own, larger tile.
struct StreamShape {
    KernShape grid, seg, seg_alt;   // seg_alt: the other form of the segmented kernel (A/B through doppler_b200_tune)
    KernShape direct;   // the lean loop with a larger tile, for launches whose samples are mostly in table-less pieces:
                        // direct evaluation is issue-bound, so the per-tile pipeline overhead is what there is to save
};

// Longest period whose de-interleaved table (period + one row + padding entries of 8 bytes) still fits
// next to the kernel's rings in the 227 KB of shared memory a CTA may have; at most kSmemTabMaxEntries.
constexpr uint32_t smem_tab_capacity(uint32_t fixed_smem, uint32_t row_samples)
{
    const uint32_t budget = 227u * 1024u, slack = row_samples + 16u;
    if (fixed_smem + (slack + 256u) * 8u >= budget) return 0;
    const uint32_t fit = (budget - fixed_smem) / 8u - slack;
    return fit < kSmemTabMaxEntries ? fit : kSmemTabMaxEntries;
}

template <int IN, int OUT, int WARPS, int S, int U, bool SEG, bool PLATEAU = false>
KernShape make_shape()
{
    using C = dmix::StreamCfg<IN, OUT, WARPS, S, U>;
    constexpr uint32_t scratch = PLATEAU ? (uint32_t)(WARPS * C::kPlateauBytes) : 0u;
    constexpr uint32_t fixed = (SEG ? (uint32_t)C::kFixedSmem : (uint32_t)C::kGridSmem) + scratch;
    MixKernel kern;
    if constexpr (SEG)   // only the kernel this shape is for gets instantiated
        kern = dmix::mix_stream_kernel<IN, OUT, WARPS, S, U>;
    else
        kern = dmix::mix_grid_kernel<IN, OUT, WARPS, S, U>;
    return KernShape{kern, WARPS, (uint32_t)C::kTileSamples, (uint32_t)C::kRow, (uint32_t)C::kGran, fixed, &C::table_bytes,
                     smem_tab_capacity(fixed, (uint32_t)C::kRow), scratch, WARPS};
}

// Q-warp pipelines (mix_pipe_kernel)
template <int IN, int OUT, int WARPS, int S, int U, int Q>
KernShape make_pipe_shape()
{
    using P = dmix::PipeCfg<IN, OUT, WARPS, S, U, Q>;
    return KernShape{dmix::mix_pipe_kernel<IN, OUT, WARPS, S, U, Q>, WARPS, (uint32_t)P::kTileSamples, (uint32_t)P::kRow, (uint32_t)P::kGran,
                     (uint32_t)P::kFixedSmem, &P::table_bytes, smem_tab_capacity((uint32_t)P::kFixedSmem, (uint32_t)P::kRow), 0u, P::kPipes};
}

// (WARPS, S, U) of the segmented kernel per type pair; the host walk of the work decomposition
// (doppler_b200_plan_tiles_trace) instantiates the same configurations.  The 4-warp pipelines of mix_pipe_kernel
// (i16 -> i16) lost the interleaved A/B against the per-warp pipelines (COLUMN 0.92 vs 0.95, cfg3 0.84 vs 0.90 of peak,
// profiles/r02_ab_seg.jsonl: the named barriers cost more than the shared issue work saves) and stay selectable
// through doppler_b200_tune for such comparisons only.
using SegI16I16 = dmix::StreamCfg<0, 0, SEG_I16I16>;
using SegI16F32 = dmix::StreamCfg<0, 1, SEG_I16F32>;
using SegF32I16 = dmix::StreamCfg<1, 0, SEG_F32I16>;
using SegF32F32 = dmix::StreamCfg<1, 1, SEG_F32F32>;

const StreamShape& shape_for(int in, int out)
{
    // lean (grid) shapes: chosen under SUSTAINED load -- launches queued back to back pull the SM clock to ~1.65 GHz
    // under the 1 kW power cap, which favours fewer, larger tiles over more warps (profiles/r01_tune_sustained.jsonl;
    // isolated launches prefer (24,2,2) for f32->i16, r01_tune_stream_smemtab_fmul2.jsonl).
    // direct shapes: profiles/r01_tune_direct_linear.jsonl
    static const StreamShape shapes[2][2] = {
        {{make_shape<0, 0, 20, 2, 2, false>(), make_shape<0, 0, SEG_I16I16, true>(), make_pipe_shape<0, 0, SEG_I16I16, 4>(),
          make_shape<0, 0, 12, 2, 6, false, true>()},
         {make_shape<0, 1, 20, 3, 2, false>(), make_shape<0, 1, SEG_I16F32, true>(), make_shape<0, 1, SEG_I16F32, true>(),
          make_shape<0, 1, 16, 2, 4, false, true>()}},
        {{make_shape<1, 0, 16, 2, 3, false>(), make_shape<1, 0, SEG_F32I16, true>(), make_shape<1, 0, SEG_F32I16, true>(),
          make_shape<1, 0, 16, 2, 6, false, true>()},
         {make_shape<1, 1, 16, 2, 2, false>(), make_shape<1, 1, SEG_F32F32, true>(), make_shape<1, 1, SEG_F32F32, true>(),
          make_shape<1, 1, 16, 2, 2, false, true>()}},
    };
    return shapes[in][out];
}

// floor-division constants for x < 2^31 (Granlund-Montgomery round-up method, N = 31)
void magic_for(uint32_t d, uint32_t* magic, uint32_t* shift)
{
    uint32_t l = 0;
    while ((1ull << l) < d) l++;
    *shift = 31 + l;
    *magic = (uint32_t)(((1ull << (31 + l)) / d) + 1);
}

// Returns the arena offset of the phasor table for (r, period), building it on `s` if needed;
// kNoTab when a table is not worthwhile.  Only periods that fit the kernel's shared-memory table
// get one; longer periods are mixed as COLUMN segments (phasors evaluated once per column and
// reused over rows) or, with fewer than kColumnMinRows whole periods, evaluated per sample.
void rt_quiesce(doppler_b200_ctx* ctx);   // the resident kernel of the per-block host path leaves (defined with it below)

int get_table(doppler_b200_ctx* ctx, float r, uint32_t period, uint64_t piece_len, cudaStream_t s, uint32_t* off_out)
{
    *off_out = dmix::kNoTab;
    if (period > kSmemTabMaxEntries) return DOPPLER_B200_OK;
    uint32_t key;
    memcpy(&key, &r, 4);
    auto it = ctx->tables.find(key);
    if (it != ctx->tables.end() && it->second.period == period) {
        *off_out = it->second.off;
        return DOPPLER_B200_OK;
    }
    if (piece_len < 2ull * period) return DOPPLER_B200_OK;   // fewer reuses than entries: evaluate directly
    const size_t entries = (size_t)period + dmix::kTabPad;
    if (!ctx->arena) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->arena, kArenaEntries * sizeof(float2)));
        ctx->arena_cap = kArenaEntries;
    }
    if (ctx->arena_used + entries > ctx->arena_cap) {
        // Recycle in place: the arena is a fixed block allocated once, so a stream that forms a new ratio per
        // block for hours (realtime track mode) never pays a free / malloc.  Launches that still read the old
        // tables are drained first -- one device-wide wait per >= 1000 table builds.
        rt_quiesce(ctx);
        CUDA_TRY(ctx, cudaDeviceSynchronize());
        ctx->arena_used = 0;
        ctx->tables.clear();
        ctx->rt_steady.valid = false;   // (it remembers a table offset)
    }
    const uint32_t off = (uint32_t)ctx->arena_used;
    const uint32_t blocks = (uint32_t)((entries + dmix::kThreads - 1) / dmix::kThreads);
    dmix::build_phasor_table_kernel<<<blocks, dmix::kThreads, 0, s>>>(ctx->arena + off, r, period, (uint32_t)entries);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    CUDA_TRY(ctx, cudaEventRecord(ctx->tables_ready, s));
    ctx->tables_event_valid = true;
    ctx->tables_stream = s;
    ctx->arena_used += (entries + 1) & ~(size_t)1;   // keep 16-byte alignment of table starts
    ctx->tables[key] = TableRef{off, period};
    *off_out = off;
    return DOPPLER_B200_OK;
}

// Cuts one launch into segments (mixer_kernels.cuh, DevSeg): COLUMN for every table-less periodic
// piece with at least kColumnMinRows whole periods, GRID everywhere else.  Returns tail_begin, the
// granule-aligned end of the segmented range.  Host-only and device-free: exercised on the CPU by
// doppler_b200_plan_tiles_trace (tests/test_plan.py).
uint32_t build_segments(const std::vector<DevPiece>& dev, uint32_t nsamp, uint32_t T, uint32_t gran, uint32_t npipes,
                        std::vector<DevSeg>* out)
{
    const uint32_t gmask = ~(gran - 1u), tail_begin = nsamp & gmask;
    std::vector<DevSeg>& segs = *out;
    // Rows per COLUMN unit.  More rows amortise the unit's window evaluation (about kWindowCostTiles tiles' worth of issue
    // slots) over more tiles; fewer, shorter units leave less work unbalanced at the end of the launch (units are claimed
    // dynamically, so the idle tail is about half a unit per pipeline).  Per-tile cost ~ kWindowCostTiles / R + R / (2 * tiles
    // per pipeline) is smallest at R = sqrt(2 * kWindowCostTiles * tiles per pipeline).  (Round 1 halved R until the launch had
    // 16 units per pipeline: R = 8 at 256 M samples, where the window evaluations were 15 % of all issued instructions --
    // profiles/r02_ncu_column_i16_i16_before.txt.)
    const double tiles_per_pipe = (double)nsamp / T / (npipes ? npipes : 1);
    uint32_t rcap0 = (uint32_t)sqrt(2.0 * kWindowCostTiles * tiles_per_pipe);
    rcap0 = std::max(kColumnMinRows, std::min(kColumnMaxRows, rcap0));
    for (uint32_t rcap = rcap0;; rcap /= 2) {
        segs.clear();
        uint32_t cursor = 0, cursor_piece = 0, units = 0;
        auto push_grid = [&](uint32_t b, uint32_t e, uint32_t piece) {
            if (e <= b) return;
            DevSeg g;
            memset(&g, 0, sizeof g);
            g.k_begin = b;
            g.k_end = e;
            g.piece = piece;
            g.unit_begin = units;
            units += ((e - b + T - 1) / T + dmix::kGridUnitTiles - 1) / dmix::kGridUnitTiles;
            g.unit_end = units;
            segs.push_back(g);
        };
        for (size_t i = 0; i < dev.size() && rcap >= kColumnMinRows && nsamp >= kColumnMinLaunch; i++) {
            const DevPiece& d = dev[i];
            if (d.period <= kSmemTabMaxEntries || d.tab != dmix::kNoTab) continue;
            const int64_t P = d.period;
            // the granule-aligned inside of the piece; its sub-granule edges stay with the GRID segments
            const int64_t kb = ((int64_t)std::max(d.k_begin, cursor) + gran - 1) & ~(int64_t)(gran - 1);
            const int64_t ke = (int64_t)(std::min(d.k_end, tail_begin) & gmask);
            if (ke - kb < (int64_t)kColumnMinRows * P) continue;
            const int64_t kph0 = (int64_t)d.k_begin + (P - d.base) % P;   // first sample of the piece with phase 0
            int64_t jlo = (kb - kph0) / P;                                // row holding kb (floor division)
            if ((kb - kph0) % P < 0) jlo--;
            const int64_t k0 = kph0 + jlo * P;                            // <= kb, may be negative
            const uint32_t rows = (uint32_t)((ke - k0 + P - 1) / P);
            push_grid(cursor, (uint32_t)kb, cursor_piece);
            DevSeg c;
            memset(&c, 0, sizeof c);
            c.k_begin = (uint32_t)kb;
            c.k_end = (uint32_t)ke;
            c.piece = (uint32_t)i;
            c.rows = rows;
            const uint32_t groups = (rows + rcap - 1) / rcap;
            c.rows_per_unit = (rows + groups - 1) / groups;
            c.ncols = (uint32_t)((P + gran - 1 + T - 1) / T);
            magic_for(c.ncols, &c.ncols_magic, &c.ncols_shift);
            c.k0 = (uint32_t)(int32_t)k0;
            c.period = d.period;
            c.r = d.r;
            c.unit_begin = units;
            units += ((c.rows + c.rows_per_unit - 1) / c.rows_per_unit) * c.ncols;
            c.unit_end = units;
            segs.push_back(c);
            cursor = (uint32_t)ke;
            cursor_piece = (uint32_t)i;
        }
        push_grid(cursor, tail_begin, cursor_piece);
        // fewer rows per unit (more, shorter units) until the pipelines can balance: units are claimed
        // dynamically, so the end-of-launch idle time is about one unit in kUnitsPerPipe
        if (units >= kUnitsPerPipe * npipes || rcap / 2 < kColumnMinRows || rcap * 2 <= rcap0) break;   // at most one halving: keep some slack for balance
    }
    return tail_begin;
}

// index[i] = segment containing work unit i << kSegIndexShift (the kernel's iterator jumps through it)
std::vector<uint32_t> build_seg_index(const std::vector<DevSeg>& segs)
{
    const uint32_t nunits = segs.empty() ? 0 : segs.back().unit_end;
    std::vector<uint32_t> index((nunits >> dmix::kSegIndexShift) + 1);
    uint32_t sgi = 0;
    for (size_t i = 0; i < index.size(); i++) {
        const uint32_t u = (uint32_t)i << dmix::kSegIndexShift;
        while (sgi + 1 < segs.size() && u >= segs[sgi].unit_end) sgi++;
        index[i] = sgi;
    }
    return index;
}

// Launch-relative device pieces of stream pieces clipped to [l0, l1) (tables not resolved: tab = kNoTab).
void clip_pieces(const std::vector<dplan::Piece>& pieces, uint64_t l0, uint64_t l1, uint32_t row_samples,
                 std::vector<DevPiece>* dev, std::vector<const dplan::Piece*>* src)
{
    for (size_t i = 0; i < pieces.size(); i++) {
        const dplan::Piece& pc = pieces[i];
        if (pc.k_end <= l0 || pc.k_begin >= l1) continue;
        const uint64_t b = std::max(pc.k_begin, l0), e = std::min(pc.k_end, l1);
        const uint64_t delta = b - pc.k_begin;
        DevPiece d;
        memset(&d, 0, sizeof d);
        d.k_begin = (uint32_t)(b - l0);
        d.k_end = (uint32_t)(e - l0);
        d.period = pc.period;
        d.r = pc.r;
        d.tab = dmix::kNoTab;
        if (pc.period == 0) {
            d.base = pc.base + (uint32_t)delta;
        } else {
            d.base = (uint32_t)(((uint64_t)pc.base + delta) % pc.period);
            magic_for(pc.period, &d.magic, &d.shift);
            d.step_u = row_samples % pc.period;
        }
        dev->push_back(d);
        if (src) src->push_back(&pc);
    }
}

// Host walk of one launch's work decomposition (doppler_b200_plan_tiles_trace).
template <typename C>
long tiles_trace(const std::vector<DevPiece>& dev, uint32_t nsamp, uint32_t npipes, uint32_t* trace, uint32_t* cover,
                 uint64_t* stats)
{
    std::vector<DevSeg> segs;
    const uint32_t tail_begin = build_segments(dev, nsamp, (uint32_t)C::kTileSamples, (uint32_t)C::kGran, npipes, &segs);
    MixArgs a;
    memset(&a, 0, sizeof a);
    a.nsamples = nsamp;
    a.npieces = (uint32_t)dev.size();
    a.pieces = dev.data();
    a.nsegs = (uint32_t)segs.size();
    a.segs = segs.data();
    const std::vector<uint32_t> index = build_seg_index(segs);
    a.seg_index = index.data();
    for (size_t i = 0; i < segs.size() && i < (size_t)dmix::kInlineSegs; i++) a.inl_segs[i] = segs[i];
    a.nunits = segs.empty() ? 0 : segs.back().unit_end;
    a.tail_begin = tail_begin;
    uint64_t ncol = 0, ntiles = 0, col_tiles = 0, col_windows = 0, col_samples = 0;
    for (const DevSeg& g : segs) ncol += g.rows != 0;
    for (uint32_t pipe = 0; pipe < npipes; pipe++) {
        dmix::TileIter<C> it;
        it.init(a, pipe, npipes);
        dmix::TileDesc d;
        size_t pi = 0;
        while (it.next(a, d)) {
            ntiles++;
            if (d.info & dmix::kColFlag) {
                col_tiles++;
                col_samples += d.nsamp;
                col_windows += (d.info & dmix::kColFirst) != 0;
            }
            if (d.nsamp == 0 || d.nsamp > (uint32_t)C::kTileSamples || (d.k0 | d.nsamp) % C::kGran) return -2;
            if (d.skip >= d.nsamp || d.skip % C::kGran) return -4;
            for (uint32_t x = d.skip; x < d.nsamp; x++) {
                const uint32_t k = d.k0 + x;
                if (k >= nsamp) return -3;
                uint32_t n;
                if (d.info & dmix::kColFlag) {
                    // the kernel's COLUMN arithmetic: window entry x + kWinLead - s_j of the window that
                    // starts at phase phase0 - kWinLead (mod period)
                    const DevSeg& g = segs[d.seg];
                    const uint32_t e = x + dmix::kWinLead - (d.info & 0xffu);
                    const uint32_t f0 = d.phase0 >= (uint32_t)dmix::kWinLead ? d.phase0 - dmix::kWinLead
                                                                              : d.phase0 + g.period - dmix::kWinLead;
                    n = (uint32_t)(((uint64_t)f0 + e) % g.period) + 1u;
                } else {
                    while (pi + 1 < dev.size() && k >= dev[pi].k_end) pi++;
                    while (pi > 0 && k < dev[pi].k_begin) pi--;
                    const DevPiece& p = dev[pi];
                    const uint32_t off = k - p.k_begin;
                    n = p.period ? (uint32_t)(((uint64_t)p.base + off) % p.period) + 1u : p.base + off;
                }
                trace[k] = n;
                cover[k]++;
            }
        }
    }
    if (stats) {
        stats[0] = segs.size();
        stats[1] = ncol;
        stats[2] = a.nunits;
        stats[3] = ntiles;
        stats[4] = col_tiles;
        stats[5] = col_windows;
        stats[6] = col_samples;
        stats[7] = (uint64_t)C::kTileSamples;
    }
    return (long)tail_begin;
}

// Enqueues the mixer over device buffers for a list of constant-shift runs.
using SmallKernel = void (*)(const MixArgs, const dmix::SmallDone);
constexpr int kSmallV = 4;   // groups per thread per step in the throughput flavour of the small kernel
SmallKernel small_kernel_for(int in, int out, bool wide)
{
    static const SmallKernel k1[2][2] = {{dmix::mix_small_kernel<0, 0, 1>, dmix::mix_small_kernel<0, 1, 1>},
                                         {dmix::mix_small_kernel<1, 0, 1>, dmix::mix_small_kernel<1, 1, 1>}};
    static const SmallKernel kv[2][2] = {{dmix::mix_small_kernel<0, 0, kSmallV>, dmix::mix_small_kernel<0, 1, kSmallV>},
                                         {dmix::mix_small_kernel<1, 0, kSmallV>, dmix::mix_small_kernel<1, 1, kSmallV>}};
    return wide ? kv[in][out] : k1[in][out];
}

// `done` (optional): have the last CTA of a small launch raise a flag in host memory (zero-copy per-block path).
// `rt_args` (optional, with `done`): if the call is ONE small launch whose pieces fit the kernel arguments, do not launch:
// hand the arguments back (*rt_filled = true) for the resident kernel's mailbox.
int launch_mix(doppler_b200_ctx* ctx, const void* d_in, void* d_out, uint64_t nsamples, int intype, int outtype,
               const std::vector<dplan::Run>& runs, uint32_t* samplenum, cudaStream_t s, const dmix::SmallDone* done = nullptr,
               MixArgs* rt_args = nullptr, bool* rt_filled = nullptr)
{
    if (nsamples == 0) return DOPPLER_B200_OK;
    std::vector<dplan::Piece> pieces;
    uint32_t sn_after = *samplenum;   // committed only when every launch has been enqueued
    ctx->planner.plan(runs, 0, &sn_after, &pieces);

    const StreamShape& shapes = shape_for(intype, outtype);
    const KernShape& seg_shape = ctx->seg_alt ? shapes.seg_alt : shapes.seg;
    const size_t ibps = bytes_per_sample(intype), obps = bytes_per_sample(outtype);
    // launches are cut at a common multiple of both kernels' tiles so that every launch but the last has no ragged tail
    const uint64_t lcm_tile = std::lcm<uint64_t>(std::lcm<uint64_t>(shapes.grid.tile_samples, std::lcm<uint64_t>(shapes.seg.tile_samples, shapes.seg_alt.tile_samples)), shapes.direct.tile_samples);
    const uint64_t launch_max = kLaunchMaxSamples / lcm_tile * lcm_tile;

    // table builds are ordered before this launch: by stream order when they were enqueued on `s` itself, through the event otherwise
    if (ctx->tables_event_valid && s != ctx->tables_stream) CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->tables_ready, 0));

    for (uint64_t l0 = 0; l0 < nsamples; l0 += launch_max) {
        const uint64_t l1 = std::min(nsamples, l0 + launch_max);
        std::vector<DevPiece> dev;
        std::vector<const dplan::Piece*> src;
        std::vector<uint64_t> dev_len;
        // DevPiece::step_u is baked from the row width, which depends on the type pair only (kRow = 32 * G): one value serves all three shapes
        static_assert(SegI16I16::kRow == 128 && SegI16F32::kRow == 64 && SegF32I16::kRow == 64 && SegF32F32::kRow == 64, "row width per type pair");
        if (shapes.grid.row_samples != shapes.seg.row_samples || shapes.grid.row_samples != shapes.direct.row_samples ||
            shapes.grid.row_samples != shapes.seg_alt.row_samples)
            return fail(ctx, DOPPLER_B200_EINVAL, "kernel shapes of one type pair disagree on the row width");
        clip_pieces(pieces, l0, l1, shapes.grid.row_samples, &dev, &src);
        for (size_t i = 0; i < dev.size(); i++) {
            dev_len.push_back(dev[i].k_end - dev[i].k_begin);
            if (dev[i].period) {
                uint64_t reuse = src[i]->k_end - src[i]->k_begin;   // samples that will read the table
                uint32_t r_bits;
                memcpy(&r_bits, &dev[i].r, 4);
                if (done && ctx->tiny_last_r_valid && r_bits == ctx->tiny_last_r_bits) reuse = std::max<uint64_t>(reuse, 2ull * dev[i].period);
                int rc = get_table(ctx, dev[i].r, dev[i].period, reuse, s, &dev[i].tab);
                if (rc) return rc;
            }
        }
        // get_table may have recycled the arena: offsets taken earlier in this launch would
        // dangle.  Re-resolve every tabled piece against the final cache (cheap, rare).
        for (size_t i = 0; i < dev.size(); i++) {
            DevPiece& d = dev[i];
            if (d.tab == dmix::kNoTab) continue;
            uint32_t key;
            memcpy(&key, &d.r, 4);
            auto it = ctx->tables.find(key);
            d.tab = (it != ctx->tables.end() && it->second.period == d.period) ? it->second.off : dmix::kNoTab;
        }

        const uint32_t nsamp = (uint32_t)(l1 - l0);
        const bool small = done != nullptr || nsamp <= ctx->small_max;   // latency-shaped kernel: no segments, no shared-memory table
        // The persistent kernels are one CTA per SM with the whole register file, and the lean one deals its tiles out
        // statically: next to a resident CTA of the per-block path (it lingers for its idle time-out after the last
        // block) one of them would wait for a neighbour to finish -- up to twice the launch time.  It leaves first.
        if (!small && ctx->rt_mb && ctx->rt_mb->alive != 0) rt_quiesce(ctx);
        std::vector<DevSeg> segs;
        uint32_t tail_begin = nsamp;
        if (!small)
            tail_begin = build_segments(dev, nsamp, seg_shape.tile_samples, seg_shape.gran, (uint32_t)ctx->sm_count * (uint32_t)seg_shape.pipes, &segs);
        // no COLUMN segment: the whole launch is one GRID segment and takes the lean loop
        const bool grid_only = small || (segs.size() <= 1 && (segs.empty() || segs[0].rows == 0));
        uint64_t tableless = 0;
        for (size_t i = 0; i < dev.size(); i++) tableless += dev[i].tab == dmix::kNoTab ? dev_len[i] : 0;
        const KernShape& shape = !grid_only ? seg_shape : 2 * tableless > nsamp ? shapes.direct : shapes.grid;
        // one table per launch is staged in shared memory: the eligible piece covering most samples
        uint32_t smem_piece = dmix::kNoPiece;
        uint64_t smem_piece_len = 0;
        for (size_t i = 0; !small && i < dev.size(); i++)
            if (dev[i].tab != dmix::kNoTab && dev[i].period <= shape.smem_tab_entries && dev_len[i] > smem_piece_len) {
                smem_piece = (uint32_t)i;
                smem_piece_len = dev_len[i];
            }

        MixArgs a;
        memset(&a, 0, sizeof a);
        a.in = static_cast<const char*>(d_in) + l0 * ibps;
        a.out = static_cast<char*>(d_out) + l0 * obps;
        a.tables = ctx->arena;
        a.nsamples = nsamp;
        a.npieces = (uint32_t)dev.size();
        a.nsegs = (uint32_t)segs.size();
        a.nunits = segs.empty() ? 0 : segs.back().unit_end;
        a.tail_begin = tail_begin;
        a.smem_piece = smem_piece;
        a.plateau_scratch = shape.scratch_smem != 0;
        a.max_claim = ctx->max_claim;
        // persistent: one CTA per SM, every warp an independent pipeline over interleaved work units
        if (grid_only) a.nunits = nsamp / shape.tile_samples;   // whole tiles; the lean loop mixes the ragged end itself
        const uint32_t want = (a.nunits + shape.pipes - 1) / shape.pipes;
        const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)ctx->sm_count, want));
        const size_t smem = shape.fixed_smem + (smem_piece != dmix::kNoPiece ? shape.table_bytes(dev[smem_piece].period) : 0);

        // Launch metadata that does not fit the kernel parameters travels as ONE image -- [work-unit counter |
        // pieces | segments | unit -> segment index] -- assembled in a pinned staging slot and sent with a single
        // asynchronous copy (a track-mode launch carries ~10^3 pieces; separate pageable copies and a memset per
        // launch cost ~10 % of a 600-piece launch).
        const bool up_pieces = dev.size() > (size_t)dmix::kInlinePieces, up_segs = segs.size() > (size_t)dmix::kInlineSegs;
        for (size_t i = 0; !up_pieces && i < dev.size(); i++) a.inl[i] = dev[i];
        for (size_t i = 0; !up_segs && i < segs.size(); i++) a.inl_segs[i] = segs[i];
        char* d_meta = nullptr;
        MetaSlot* meta_slot = nullptr;
        if (!grid_only || up_pieces) {   // (a small launch uploads only a long piece list)
            const std::vector<uint32_t> index = up_segs ? build_seg_index(segs) : std::vector<uint32_t>();
            const size_t off_pieces = 16, off_segs = off_pieces + (up_pieces ? dev.size() * sizeof(DevPiece) : 0);
            const size_t off_index = off_segs + (up_segs ? segs.size() * sizeof(DevSeg) : 0);
            const size_t bytes = off_index + index.size() * sizeof(uint32_t);
            MetaSlot& ms = ctx->meta[ctx->meta_next++ % kMetaSlots];
            if (ms.used) CUDA_TRY(ctx, cudaEventSynchronize(ms.done));   // the launch that last used this slot has finished
            if (ms.cap < bytes) {
                if (ms.host) cudaFreeHost(ms.host);
                if (ms.dev) cudaFree(ms.dev);
                ms.host = ms.dev = nullptr;
                ms.cap = 0;
                const size_t cap = std::max<size_t>(bytes * 2, 1u << 16);
                CUDA_TRY(ctx, cudaMallocHost(&ms.host, cap));
                CUDA_TRY(ctx, cudaMalloc(&ms.dev, cap));
                ms.cap = cap;
            }
            if (!ms.done) {
                CUDA_TRY(ctx, cudaEventCreateWithFlags(&ms.done, cudaEventDisableTiming));
                CUDA_TRY(ctx, cudaEventCreateWithFlags(&ms.copied, cudaEventDisableTiming));
            }
            char* h = static_cast<char*>(ms.host);
            memset(h, 0, 16);   // the work-unit counter starts at 0
            if (up_pieces) memcpy(h + off_pieces, dev.data(), dev.size() * sizeof(DevPiece));
            if (up_segs) memcpy(h + off_segs, segs.data(), segs.size() * sizeof(DevSeg));
            if (!index.empty()) memcpy(h + off_index, index.data(), index.size() * sizeof(uint32_t));
            d_meta = static_cast<char*>(ms.dev);
            // on the context's own copy stream: the upload overlaps whatever the caller's stream is still running
            // (the previous launch, when calls are queued back to back); the kernel waits for it through an event
            CUDA_TRY(ctx, cudaMemcpyAsync(d_meta, h, bytes, cudaMemcpyHostToDevice, ctx->meta_stream));
            CUDA_TRY(ctx, cudaEventRecord(ms.copied, ctx->meta_stream));
            CUDA_TRY(ctx, cudaStreamWaitEvent(s, ms.copied, 0));
            meta_slot = &ms;
            a.unit_counter = reinterpret_cast<uint32_t*>(d_meta);
            if (up_pieces) a.pieces = reinterpret_cast<const DevPiece*>(d_meta + off_pieces);
            if (up_segs) {
                a.segs = reinterpret_cast<const DevSeg*>(d_meta + off_segs);
                a.seg_index = reinterpret_cast<const uint32_t*>(d_meta + off_index);
            }
        }
        if (small && rt_args && a.npieces <= (uint32_t)dmix::kRtPieces && l0 == 0 && l1 == nsamples) {
            *rt_args = a;
            *rt_filled = true;
            *samplenum = sn_after;
            return DOPPLER_B200_OK;
        }
        if (small) {
            // one group per thread while that still fills the chip once over (latency), kSmallV per thread beyond (bytes in
            // flight); a flagged zero-copy launch of up to 1024 groups is ONE CTA: no cross-CTA counter before the flag
            const uint32_t groups = std::max<uint32_t>(1, nsamp / (uint32_t)dmix::group_samples(intype, outtype));
            const bool one_cta = done != nullptr && groups <= (uint32_t)dmix::kSmallThreads * kSmallV;
            const bool wide = one_cta || groups > (uint32_t)ctx->sm_count * 8u * dmix::kSmallThreads;
            const uint32_t per_cta = (uint32_t)dmix::kSmallThreads * (wide ? kSmallV : 1);
            const uint32_t ctas = one_cta ? 1u : std::min<uint32_t>((groups + per_cta - 1) / per_cta, (uint32_t)ctx->sm_count * 8u);   // one resident wave
            small_kernel_for(intype, outtype, wide)<<<ctas, dmix::kSmallThreads, 0, s>>>(a, done ? *done : dmix::SmallDone{nullptr, nullptr, 0});
        } else {
            const void* kfn = reinterpret_cast<const void*>(shape.kern);
            if (std::find(ctx->configured.begin(), ctx->configured.end(), kfn) == ctx->configured.end()) {
                CUDA_TRY(ctx, cudaFuncSetAttribute(shape.kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(shape.fixed_smem + shape.table_bytes(shape.smem_tab_entries))));
                ctx->configured.push_back(kfn);
            }
            shape.kern<<<grid, shape.warps * 32, smem, s>>>(a);
        }
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
        if (meta_slot) {
            CUDA_TRY(ctx, cudaEventRecord(meta_slot->done, s));
            meta_slot->used = true;
        }
    }
    *samplenum = sn_after;
    return DOPPLER_B200_OK;
}

int check_common(doppler_b200_ctx* ctx, const void* in, size_t in_len, int intype, int outtype, uint32_t* samplenum,
                 void* out, size_t out_cap, uint64_t* nsamples)
{
    if (!ctx) return DOPPLER_B200_EINVAL;
    if (!valid_type(intype) || !valid_type(outtype)) return fail(ctx, DOPPLER_B200_EINVAL, "unknown IQ data type");
    if (!samplenum) return fail(ctx, DOPPLER_B200_EINVAL, "samplenum is NULL");
    const size_t ibps = bytes_per_sample(intype);
    if (in_len % ibps != 0)
        return fail(ctx, DOPPLER_B200_EALIGN, "input length %zu is not a multiple of %zu (dsp.rs assert)", in_len, ibps);
    *nsamples = in_len / ibps;
    if (*nsamples && (!in || !out)) return fail(ctx, DOPPLER_B200_EINVAL, "NULL buffer");
    if (*nsamples * bytes_per_sample(outtype) > out_cap)
        return fail(ctx, DOPPLER_B200_ECAP, "output capacity %zu < %llu bytes needed", out_cap,
                    (unsigned long long)(*nsamples * bytes_per_sample(outtype)));
    return DOPPLER_B200_OK;
}

int ensure_slot(doppler_b200_ctx* ctx, Slot& sl, size_t in_bytes, size_t out_bytes)
{
    if (!sl.stream) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    if (sl.in_cap < in_bytes || sl.out_cap < out_bytes) rt_quiesce(ctx);   // (freeing synchronises the device)
    if (sl.in_cap < in_bytes) {
        if (sl.d_in) cudaFree(sl.d_in);
        if (sl.h_in) cudaFreeHost(sl.h_in);
        sl.d_in = sl.h_in = nullptr;
        sl.in_cap = 0;
        CUDA_TRY(ctx, cudaMalloc(&sl.d_in, in_bytes));
        CUDA_TRY(ctx, cudaMallocHost(&sl.h_in, in_bytes));
        CUDA_TRY(ctx, cudaHostGetDevicePointer(&sl.h_in_dev, sl.h_in, 0));
        sl.in_cap = in_bytes;
    }
    if (sl.out_cap < out_bytes) {
        if (sl.d_out) cudaFree(sl.d_out);
        if (sl.h_out) cudaFreeHost(sl.h_out);
        sl.d_out = sl.h_out = nullptr;
        sl.out_cap = 0;
        CUDA_TRY(ctx, cudaMalloc(&sl.d_out, out_bytes));
        CUDA_TRY(ctx, cudaMallocHost(&sl.h_out, out_bytes));
        CUDA_TRY(ctx, cudaHostGetDevicePointer(&sl.h_out_dev, sl.h_out, 0));
        sl.out_cap = out_bytes;
    }
    return DOPPLER_B200_OK;
}

bool is_pinned(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// Staging copies between the caller's pageable buffers and the pinned slots.  One thread moves ~10 GB/s, a fifth of what
// PCIe takes, so large copies are spread over a small pool of persistent worker threads (created on the first large copy;
// round 1 spawned up to 4 threads per copy: 1.0-1.5 Gsample/s end to end, profiles/r01_percall_latency.jsonl).  Callers
// with pinned buffers (doppler_b200_host_alloc / _register) never come here.
class CopyPool {
public:
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            quit_ = true;
        }
        cv_work_.notify_all();
        for (std::thread& t : workers_) t.join();
    }

    void copy(void* dst, const void* src, size_t bytes)
    {
        constexpr size_t kPerThread = 2u << 20;
        const unsigned hw = std::thread::hardware_concurrency();
        const size_t want = std::min<size_t>({bytes / kPerThread, (size_t)(hw > 2 ? hw / 2 : 1), (size_t)kMaxThreads});
        if (want <= 1) {
            memcpy(dst, src, bytes);
            return;
        }
        while (workers_.size() + 1 < want) {   // (no job is in flight here: copy() returns only when all workers are done)
            const size_t id = workers_.size();
            const uint64_t gen = generation_;   // a new worker must not mistake an earlier job for a fresh one
            workers_.emplace_back([this, id, gen] { run(id, gen); });
        }
        const size_t parts = workers_.size() + 1;
        const size_t part = (bytes / parts + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> g(m_);
            dst_ = static_cast<char*>(dst);
            src_ = static_cast<const char*>(src);
            bytes_ = bytes;
            part_ = part;
            pending_ = (int)workers_.size();
            generation_++;
        }
        cv_work_.notify_all();
        memcpy(dst, src, std::min(part, bytes));   // the caller takes part 0
        std::unique_lock<std::mutex> g(m_);
        cv_done_.wait(g, [&] { return pending_ == 0; });
    }

private:
    static constexpr int kMaxThreads = 8;
    void run(size_t id, uint64_t seen)
    {
        for (;;) {
            char* dst;
            const char* src;
            size_t bytes, part;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_work_.wait(g, [&] { return quit_ || generation_ != seen; });
                if (quit_) return;
                seen = generation_;
                dst = dst_, src = src_, bytes = bytes_, part = part_;
            }
            const size_t off = (id + 1) * part;
            if (off < bytes) memcpy(dst + off, src + off, std::min(part, bytes - off));
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> workers_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, part_ = 0;
    int pending_ = 0;
    uint64_t generation_ = 0;
    bool quit_ = false;
};

void staged_copy(doppler_b200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx->copy_pool) ctx->copy_pool = new CopyPool;
    static_cast<CopyPool*>(ctx->copy_pool)->copy(dst, src, bytes);
}

int retire_slot(doppler_b200_ctx* ctx, Slot& sl)
{
    if (!sl.busy) return DOPPLER_B200_OK;
    CUDA_TRY(ctx, cudaEventSynchronize(sl.done));
    if (sl.user_out) staged_copy(ctx, sl.user_out, sl.h_out, sl.user_out_bytes);
    sl.user_out = nullptr;
    sl.busy = false;
    return DOPPLER_B200_OK;
}

// Device-usable alias of a pinned host pointer (identical under unified addressing; asked for, not assumed).
int device_alias(doppler_b200_ctx* ctx, const void* host, void** dev)
{
    CUDA_TRY(ctx, cudaHostGetDevicePointer(dev, const_cast<void*>(host), 0));
    return DOPPLER_B200_OK;
}

// ---- resident kernel (mixer_kernels.cuh: mix_resident_kernel) --------------------------------------------------------------
static_assert(kTinyStageBytes == (size_t)dmix::kRtStageBytes, "the resident kernel stages a whole block in shared memory");
constexpr size_t kRtOutUnits = 2 * kTinyStageBytes / 4;   // result words of the largest block served (i16 -> f32 doubles the bytes)

// A kernel that reads blocks at `in_dev` and tables at `tables_dev`, and has served every request up to ctx->rt_seq (`pending`:
// up to the one before -- the mailbox already holds ctx->rt_seq for it).
int rt_start(doppler_b200_ctx* ctx, const void* in_dev, const void* tables_dev, bool pending)
{
    if (!ctx->rt_mb) {
        CUDA_TRY(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->rt_mb), sizeof(dmix::RtMailbox), cudaHostAllocMapped));
        memset(ctx->rt_mb, 0, sizeof(dmix::RtMailbox));
        void* alias = nullptr;
        CUDA_TRY(ctx, cudaHostGetDevicePointer(&alias, ctx->rt_mb, 0));
        ctx->rt_mb_dev = static_cast<dmix::RtMailbox*>(alias);
        CUDA_TRY(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->rt_out), kRtOutUnits * 8, cudaHostAllocMapped));
        memset(ctx->rt_out, 0, kRtOutUnits * 8);   // (no unit carries a request number yet: they start at 1)
        CUDA_TRY(ctx, cudaHostGetDevicePointer(&ctx->rt_out_dev, ctx->rt_out, 0));
        // highest priority: when the chip is full of another stream's CTAs, the one resident CTA is placed first
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->rt_stream, cudaStreamNonBlocking, prio_hi));
    }
    const uint32_t gen = ++ctx->rt_gen ? ctx->rt_gen : ++ctx->rt_gen;   // never 0
    uint32_t last = ctx->rt_seq;
    if (pending) last = (last - 1u) ? last - 1u : last - 2u;   // (request numbers skip 0)
    ctx->rt_mb->alive = gen;
    ctx->rt_in_dev = in_dev;
    ctx->rt_tables_dev = tables_dev;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    // (queued behind a predecessor that is still leaving: same stream)
    dmix::mix_resident_kernel<<<1, dmix::kSmallThreads, 0, ctx->rt_stream>>>(ctx->rt_mb_dev, in_dev, ctx->rt_out_dev, static_cast<const float2*>(tables_dev),
                                                                            ctx->rt_idle_us * 1000ull, gen, last);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->rt_starts++;
    return DOPPLER_B200_OK;
}

// A request in the mailbox's tagged sectors: payload first, then the sector's tag (mixer_kernels.cuh: RtMailbox; collect.cpp: post).
static uint32_t rt_post(doppler_b200_ctx* ctx, uint32_t head, const MixArgs* a)
{
    uint32_t payload[dmix::kRtPayloadWords] = {0};
    payload[0] = head;
    if (a) {
        payload[1] = a->nsamples;
        for (uint32_t p = 0; p < a->npieces; p++) memcpy(payload + 2 + p * dmix::kRtPieceWords, &a->inl[p], dmix::kRtPieceWords * 4);
    }
    uint32_t seq = ++ctx->rt_seq;
    if (seq == 0) {
        // the numbers wrap: no unit may still carry one that is about to be used again (the kernel is idle: every request
        // before this one has been collected, and nothing writes units between requests)
        memset(ctx->rt_out, 0, kRtOutUnits * 8);
        seq = ++ctx->rt_seq;   // (0 is the mailbox's initial state)
    }
    static_assert(offsetof(dmix::RtMailbox, req) == 0 && sizeof(dmix::RtSector) == 32, "the request sectors lead the mailbox");
    dcollect::post(reinterpret_cast<volatile uint32_t*>(ctx->rt_mb), dmix::kRtSectors, seq, payload, dmix::kRtPayloadWords);
    return seq;
}

// The resident kernel leaves now (before anything that synchronises the device, frees memory it may read, or ends the context).
void rt_quiesce(doppler_b200_ctx* ctx)
{
    if (!ctx->rt_mb) return;
    if (ctx->rt_mb->alive != 0) rt_post(ctx, dmix::kRtQuit, nullptr);
    cudaStreamSynchronize(ctx->rt_stream);
    ctx->rt_mb->alive = 0;
    cudaGetLastError();
}

// One request through the mailbox; returns when the `out_words` words of its result are in `out`.
int rt_request(doppler_b200_ctx* ctx, const MixArgs& a, int intype, int outtype, void* out, size_t out_words)
{
    if (ctx->rt_mb && ctx->rt_mb->alive != 0 && (a.in != ctx->rt_in_dev || a.tables != ctx->rt_tables_dev)) rt_quiesce(ctx);
    if (!ctx->rt_mb || ctx->rt_mb->alive == 0) {
        int rc = rt_start(ctx, a.in, a.tables, false);
        if (rc) return rc;
    }
    const uint32_t seq = rt_post(ctx, dmix::rt_head(intype, outtype, a.npieces), &a);
    ctx->rt_requests++;
    unsigned char* dst = static_cast<unsigned char*>(out);
    size_t got = 0;
    for (uint32_t spins = 1;; spins++) {
        const size_t now = dcollect::collect(ctx->rt_out, seq, dst, got, out_words);
        if (now == out_words) break;
        if (now != got) spins = 1;
        got = now;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0x3fffu) == 0) {   // nothing new for ~100 us: did the kernel leave (idle time-out racing this request) or die?
            const cudaError_t q = cudaStreamQuery(ctx->rt_stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) return fail(ctx, DOPPLER_B200_ECUDA, "resident kernel failed: %s", cudaGetErrorString(q));
            if (q == cudaSuccess && dcollect::collect(ctx->rt_out, seq, dst, got, out_words) != out_words) {
                // gone without having seen the request (a kernel that saw it serves it before it leaves)
                int rc = rt_start(ctx, a.in, a.tables, true);
                if (rc) return rc;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return DOPPLER_B200_OK;
}

int tiny_host_call(doppler_b200_ctx* ctx, const void* in, uint64_t nsamples, int intype, int outtype, const float* shifts,
                   size_t nblocks, uint64_t block_samples, uint32_t samplerate, uint32_t* samplenum, void* out)
{
    const size_t ibps = bytes_per_sample(intype), obps = bytes_per_sample(outtype);
    Slot& sl = ctx->slots[0];
    int rc = retire_slot(ctx, sl);
    if (rc) return rc;
    rc = ensure_slot(ctx, sl, ctx->tiny_host_bytes, ctx->tiny_host_bytes * 2);
    if (rc) return rc;
    if (!ctx->done_flag) {
        CUDA_TRY(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->done_flag), 64, cudaHostAllocMapped));
        *ctx->done_flag = 0;
        void* alias = nullptr;
        rc = device_alias(ctx, ctx->done_flag, &alias);
        if (rc) return rc;
        ctx->done_flag_dev = static_cast<uint32_t*>(alias);
        CUDA_TRY(ctx, cudaMalloc(&ctx->done_counter, 4));
        CUDA_TRY(ctx, cudaMemset(ctx->done_counter, 0, 4));
    }
    // Up to kTinyStageBytes the block is simply copied through the slot's pinned staging (an 8 KiB memcpy costs less than
    // asking the driver what kind of memory the caller's pointers are); larger tiny calls use the caller's buffers in place
    // when they are pinned, mapped and 16-byte aligned.
    void *src_dev = sl.h_in_dev, *dst_dev = sl.h_out_dev;
    void* dst = sl.h_out;
    auto alias_ok = [&](const void* host, void** dev) {
        if (!is_pinned(host)) return false;
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, const_cast<void*>(host), 0) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        if ((uintptr_t)d & 15) return false;
        *dev = d;
        return true;
    };
    const auto t_0 = std::chrono::steady_clock::now();
    const bool probe = nsamples * ibps > kTinyStageBytes;
    if (!(probe && alias_ok(in, &src_dev))) memcpy(sl.h_in, in, nsamples * ibps);
    if (probe && alias_ok(out, &dst_dev)) dst = out;
    const auto t_1 = std::chrono::steady_clock::now();
    const bool rt_try = ctx->rt_idle_us != 0 && !probe;
    if (rt_try && ctx->rt_steady.valid && (nblocks <= 1 || block_samples == 0)) {
        // Steady state of a block stream (the reference's pump, main.rs:62-99): the same ratio as the last block and samplenum
        // inside its period.  The planner's rule for that state (plan.cpp, Planner::plan: one periodic piece, base =
        // samplenum - 1, samplenum' = (samplenum - 1 + n) mod P + 1) is applied here directly, with the table, period and
        // division constants the last block's plan carried.
        const float r = dplan::ratio(shifts[0], samplerate);
        uint32_t r_bits;
        memcpy(&r_bits, &r, 4);
        const DevPiece& last = ctx->rt_steady.piece;
        const uint32_t sn = *samplenum;
        if (r_bits == ctx->rt_steady.r_bits && sn >= 1 && sn <= last.period) {
            MixArgs a;
            memset(&a, 0, sizeof a);
            a.in = src_dev;
            a.tables = ctx->arena;
            a.nsamples = (uint32_t)nsamples;
            a.npieces = 1;
            a.inl[0] = last;
            a.inl[0].k_begin = 0;
            a.inl[0].k_end = (uint32_t)nsamples;
            a.inl[0].base = sn - 1u;
            const auto t_2r = std::chrono::steady_clock::now();
            rc = rt_request(ctx, a, intype, outtype, out, nsamples * obps / 4);
            if (rc) return rc;
            *samplenum = (uint32_t)(((uint64_t)(sn - 1u) + nsamples) % last.period) + 1u;
            if (ctx->trace && ctx->rt_requests > 64) {
                auto ns = [](auto a, auto b) { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
                ctx->rt_steady_traced++;
                ctx->rt_steady_ns[0] += ns(t_0, t_1), ctx->rt_steady_ns[1] += ns(t_1, t_2r), ctx->rt_steady_ns[2] += ns(t_2r, std::chrono::steady_clock::now());
            }
            return DOPPLER_B200_OK;
        }
    }
    std::vector<dplan::Run> runs;
    if (nblocks <= 1 || block_samples == 0)
        runs.push_back(dplan::Run{nsamples, dplan::ratio(shifts[0], samplerate)});
    else
        runs = dplan::runs_from_blocks(shifts, nblocks, block_samples, samplerate, nsamples);
    struct NoteRatio {   // on every way out: what ratio this call had (one shift), for the next call's table decision
        doppler_b200_ctx* ctx;
        const std::vector<dplan::Run>& runs;
        ~NoteRatio()
        {
            ctx->tiny_last_r_valid = runs.size() == 1;
            if (ctx->tiny_last_r_valid) memcpy(&ctx->tiny_last_r_bits, &runs[0].r, 4);
        }
    } note_ratio{ctx, runs};
    const uint32_t token = ++ctx->done_token ? ctx->done_token : ++ctx->done_token;   // never 0
    const dmix::SmallDone done{ctx->done_counter, ctx->done_flag_dev, token};
    // staged blocks (the reference's 8192-byte pump block) go to the resident kernel when their plan fits its mailbox
    MixArgs rt_args;
    bool rt_filled = false;
    const uint64_t launches_before = ctx->launches;
    rc = launch_mix(ctx, src_dev, dst_dev, nsamples, intype, outtype, runs, samplenum, sl.stream, &done, rt_try ? &rt_args : nullptr, &rt_filled);
    if (rc) return rc;
    if (rt_filled) {
        // a phasor table this block needs may have been enqueued just now, or earlier on another stream
        if (ctx->launches != launches_before) CUDA_TRY(ctx, cudaStreamSynchronize(sl.stream));
        if (ctx->tables_event_valid) {
            CUDA_TRY(ctx, cudaEventSynchronize(ctx->tables_ready));
            ctx->tables_event_valid = false;
        }
        const auto t_2r = std::chrono::steady_clock::now();
        rc = rt_request(ctx, rt_args, intype, outtype, out, nsamples * obps / 4);   // (collects the result straight into `out`)
        if (rc) return rc;
        // (a piece that could have a table but has none yet -- the first block of a ratio -- goes through launch_mix once more,
        //  which builds the table on second sight)
        ctx->rt_steady.valid = !ctx->rt_steady.off && runs.size() == 1 && rt_args.npieces == 1 && rt_args.inl[0].period != 0 &&
                               (rt_args.inl[0].tab != dmix::kNoTab || rt_args.inl[0].period > kSmemTabMaxEntries);
        if (ctx->rt_steady.valid) {
            ctx->rt_steady.piece = rt_args.inl[0];
            memcpy(&ctx->rt_steady.r_bits, &rt_args.inl[0].r, 4);
        }
        const auto t_3r = std::chrono::steady_clock::now();
        if (ctx->trace) {
            auto ns = [](auto a, auto b) { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
            if (ctx->rt_requests > 64) {   // (the first calls start the kernel, build tables, allocate: not the steady state)
                ctx->rt_traced++;
                ctx->rt_ns[0] += ns(t_0, t_1), ctx->rt_ns[1] += ns(t_1, t_2r), ctx->rt_ns[2] += ns(t_2r, t_3r);
            }
        }
        return DOPPLER_B200_OK;
    }
    const auto t_2 = std::chrono::steady_clock::now();
    // spin on the flag (host memory, written by the kernel's last CTA after a system-wide fence); a launch that died never
    // raises it, so the stream is consulted now and then
    volatile uint32_t* flag = ctx->done_flag;
    for (uint32_t spins = 0; *flag != token; spins++) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0xfffffu) == 0xfffffu) {
            const cudaError_t q = cudaStreamQuery(sl.stream);
            if (q == cudaSuccess) break;   // finished (the flag store is visible by now or the kernel did not run at all)
            if (q != cudaErrorNotReady) return fail(ctx, DOPPLER_B200_ECUDA, "small launch failed: %s", cudaGetErrorString(q));
        }
    }
    if (*flag != token) {
        CUDA_TRY(ctx, cudaStreamSynchronize(sl.stream));
        if (*flag != token) return fail(ctx, DOPPLER_B200_ECUDA, "small launch finished without raising its completion flag");
    }
    if (ctx->tables_stream == sl.stream) ctx->tables_event_valid = false;   // everything enqueued on this stream has finished
    const auto t_3 = std::chrono::steady_clock::now();
    if (dst != out) memcpy(out, dst, nsamples * obps);
    if (ctx->trace) {
        const auto t_4 = std::chrono::steady_clock::now();
        auto ns = [](auto a, auto b) { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
        ctx->tiny_calls++;
        ctx->tiny_ns[0] += ns(t_0, t_1), ctx->tiny_ns[1] += ns(t_1, t_2), ctx->tiny_ns[2] += ns(t_2, t_3), ctx->tiny_ns[3] += ns(t_3, t_4);
    }
    return DOPPLER_B200_OK;
}

// After a failed host-path call: nothing may stay in flight into the caller's buffers, and no slot may keep a
// pending copy-out into memory the caller is free to release once the call has returned.
void abandon_slots(doppler_b200_ctx* ctx)
{
    for (Slot& sl : ctx->slots) {
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        sl.user_out = nullptr;
        sl.user_out_bytes = 0;
        sl.busy = false;
    }
    cudaGetLastError();
}

// Host-buffer pipeline: chunk c uses slot c % kSlots; each slot has its own stream so the H2D
// of chunk c+1 overlaps the kernel of chunk c and the D2H of chunk c-1.  `copy_only` skips the
// kernel (doppler_b200_pipeline_probe: the ceiling the platform's host memory / PCIe path sets).
int mix_host_run(doppler_b200_ctx* ctx, const void* in, uint64_t nsamples, int intype, int outtype,
                 const float* shifts, size_t nblocks, uint64_t block_samples, uint32_t samplerate, uint32_t* samplenum,
                 void* out, bool copy_only)
{
    const size_t ibps = bytes_per_sample(intype), obps = bytes_per_sample(outtype);
    // chunk = whole number of shift blocks and of the largest tile
    // DOPPLER_B200_CHUNK_MB: tuning knob for the pipeline chunk (tools/tune); the default was chosen on B200
    static const size_t chunk_bytes = [] {
        const char* e = getenv("DOPPLER_B200_CHUNK_MB");
        const long mb = e ? atol(e) : 0;
        return mb >= 1 && mb <= 1024 ? (size_t)mb << 20 : kHostChunkBytes;
    }();
    uint64_t chunk = chunk_bytes / ibps;
    if (block_samples && nblocks > 1) chunk = std::max<uint64_t>(block_samples, chunk / block_samples * block_samples);
    uint32_t sn = *samplenum;
    if (!copy_only && nsamples * ibps <= ctx->tiny_host_bytes) {
        // The reference's own call granularity (one 8192-byte pump block per shift_frequency call, main.rs:49,70): two
        // cudaMemcpyAsync + an event wait cost more than the work.  Zero-copy instead: the latency-shaped kernel reads the
        // block from mapped pinned host memory and writes the result there, and the host waits on a flag the kernel's last
        // CTA raises in host memory.
        int rc = tiny_host_call(ctx, in, nsamples, intype, outtype, shifts, nblocks, block_samples, samplerate, &sn, out);
        if (rc) return rc;
        *samplenum = sn;
        return DOPPLER_B200_OK;
    }
    const bool in_pinned = is_pinned(in), out_pinned = is_pinned(out);
    int c = 0;
    for (uint64_t k = 0; k < nsamples; k += chunk, c++) {
        const uint64_t n = std::min(chunk, nsamples - k);
        Slot& sl = ctx->slots[c % kSlots];
        int rc = retire_slot(ctx, sl);
        if (rc) return rc;
        rc = ensure_slot(ctx, sl, std::min<uint64_t>(chunk, nsamples) * ibps, std::min<uint64_t>(chunk, nsamples) * obps);
        if (rc) return rc;
        const char* src = static_cast<const char*>(in) + k * ibps;
        if (!in_pinned) {
            staged_copy(ctx, sl.h_in, src, n * ibps);
            src = static_cast<const char*>(sl.h_in);
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(sl.d_in, src, n * ibps, cudaMemcpyHostToDevice, sl.stream));
        if (!copy_only) {
            std::vector<dplan::Run> runs;
            if (nblocks <= 1 || block_samples == 0) {
                runs.push_back(dplan::Run{n, dplan::ratio(shifts[0], samplerate)});
            } else {
                const size_t b0 = (size_t)(k / block_samples);
                runs = dplan::runs_from_blocks(shifts + b0, nblocks - b0, block_samples, samplerate, n);
            }
            rc = launch_mix(ctx, sl.d_in, sl.d_out, n, intype, outtype, runs, &sn, sl.stream);
            if (rc) return rc;
        }
        char* dst = static_cast<char*>(out) + k * obps;
        if (out_pinned) {
            CUDA_TRY(ctx, cudaMemcpyAsync(dst, sl.d_out, n * obps, cudaMemcpyDeviceToHost, sl.stream));
            sl.user_out = nullptr;
        } else {
            CUDA_TRY(ctx, cudaMemcpyAsync(sl.h_out, sl.d_out, n * obps, cudaMemcpyDeviceToHost, sl.stream));
            sl.user_out = dst;
            sl.user_out_bytes = n * obps;
        }
        CUDA_TRY(ctx, cudaEventRecord(sl.done, sl.stream));
        sl.busy = true;
    }
    for (int i = 0; i < kSlots; i++) {
        int rc = retire_slot(ctx, ctx->slots[i]);
        if (rc) return rc;
    }
    *samplenum = sn;
    return DOPPLER_B200_OK;
}

int mix_host(doppler_b200_ctx* ctx, const void* in, uint64_t nsamples, int intype, int outtype, const float* shifts,
             size_t nblocks, uint64_t block_samples, uint32_t samplerate, uint32_t* samplenum, void* out, bool copy_only = false)
{
    const int rc = mix_host_run(ctx, in, nsamples, intype, outtype, shifts, nblocks, block_samples, samplerate, samplenum, out, copy_only);
    if (rc) {
        const std::string keep = ctx->err;   // the first error is the one to report
        abandon_slots(ctx);
        ctx->err = keep;
    }
    return rc;
}

int check_blocks(doppler_b200_ctx* ctx, size_t in_len, int intype, const float* shifts, size_t nblocks, size_t block_bytes,
                 uint64_t* block_samples)
{
    if (!shifts || nblocks == 0) return fail(ctx, DOPPLER_B200_EINVAL, "no shift schedule");
    const size_t ibps = bytes_per_sample(intype);
    if (block_bytes == 0 || block_bytes % ibps != 0)
        return fail(ctx, DOPPLER_B200_EINVAL, "block_bytes %zu is not a whole number of samples", block_bytes);
    if ((in_len + block_bytes - 1) / block_bytes > nblocks)
        return fail(ctx, DOPPLER_B200_EINVAL, "shift schedule has %zu blocks, input needs %zu", nblocks,
                    (in_len + block_bytes - 1) / block_bytes);
    *block_samples = block_bytes / ibps;
    return DOPPLER_B200_OK;
}

}  // namespace

// ===============================================================================================
extern "C" {

int doppler_b200_abi_version(void) { return DOPPLER_B200_ABI_VERSION; }

int doppler_b200_create(int device, doppler_b200_ctx** ctx_out)
{
    if (!ctx_out) return DOPPLER_B200_EINVAL;
    *ctx_out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, DOPPLER_B200_ENODEV, "no CUDA device: %s (this library has no CPU path)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(nullptr, DOPPLER_B200_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop;
    CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, DOPPLER_B200_ENODEV, "device %d is sm_%d%d; this build carries sm_100a code only", device,
                    prop.major, prop.minor);
    if (!doppler_b200_libm_compatible()) {
        // numerics contract: the device reproduces glibc's __sincosf_fma; this host's libm differs
        if (getenv("DOPPLER_B200_STRICT_LIBM") && atoi(getenv("DOPPLER_B200_STRICT_LIBM")) != 0)
            return fail(nullptr, DOPPLER_B200_ELIBM, "host libm sincosf differs from the glibc __sincosf_fma sequence the device reproduces");
        static bool warned = false;
        if (!warned) {
            warned = true;
            fprintf(stderr, "doppler_b200: warning: this host's libm sincosf is not bit-identical to glibc x86-64 __sincosf_fma; GPU output "
                            "matches that variant, not the reference built on this host (doppler_b200_libm_compatible() == 0)\n");
        }
    }
    CUDA_TRY(nullptr, cudaSetDevice(device));
    doppler_b200_ctx* ctx = new (std::nothrow) doppler_b200_ctx;
    if (!ctx) return fail(nullptr, DOPPLER_B200_ENOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->trace = getenv("DOPPLER_B200_TRACE") != nullptr;
    if (const char* e = getenv("DOPPLER_B200_SMALL_MAX")) ctx->small_max = (uint32_t)strtoul(e, nullptr, 10);   // tuning knobs (tools/tune)
    if (const char* e = getenv("DOPPLER_B200_TINY_BYTES")) ctx->tiny_host_bytes = (size_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("DOPPLER_B200_RESIDENT_IDLE_US")) ctx->rt_idle_us = strtoull(e, nullptr, 10);
    if (const char* e = getenv("DOPPLER_B200_NO_STEADY_RULE")) ctx->rt_steady.off = atoi(e) != 0;
    cudaError_t e2 = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&ctx->meta_stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ctx->tables_ready, cudaEventDisableTiming);
    // (the kernels' shared-memory attributes are set on first use, launch_mix: configuring all of them here loads every
    //  kernel of the module, which a CLI run over a second of IQ never needs)
    if (e2 != cudaSuccess) {
        fail(nullptr, DOPPLER_B200_ECUDA, "context setup failed: %s", cudaGetErrorString(e2));
        delete ctx;
        return DOPPLER_B200_ECUDA;
    }
    *ctx_out = ctx;
    return DOPPLER_B200_OK;
}

void doppler_b200_destroy(doppler_b200_ctx* ctx)
{
    if (!ctx) return;
    if (ctx->trace && ctx->tiny_calls)
        fprintf(stderr, "{\"tiny_host_calls\": %llu, \"ns_per_call\": {\"stage_in\": %.0f, \"plan_and_launch\": %.0f, \"wait_flag\": %.0f, \"copy_out\": %.0f}}\n",
                (unsigned long long)ctx->tiny_calls, (double)ctx->tiny_ns[0] / ctx->tiny_calls, (double)ctx->tiny_ns[1] / ctx->tiny_calls,
                (double)ctx->tiny_ns[2] / ctx->tiny_calls, (double)ctx->tiny_ns[3] / ctx->tiny_calls);
    if (ctx->trace && ctx->rt_traced)
        fprintf(stderr, "{\"resident_kernel_calls\": %llu, \"kernel_starts\": %llu, \"ns_per_call_after_the_first_64\": {\"stage_in\": %.0f, \"plan\": %.0f, \"request_to_collected\": %.0f}}\n",
                (unsigned long long)ctx->rt_requests, (unsigned long long)ctx->rt_starts, (double)ctx->rt_ns[0] / ctx->rt_traced,
                (double)ctx->rt_ns[1] / ctx->rt_traced, (double)ctx->rt_ns[2] / ctx->rt_traced);
    if (ctx->trace && ctx->rt_steady_traced)
        fprintf(stderr, "{\"resident_kernel_calls_planned_by_the_steady_state_rule\": %llu, \"ns_per_call\": {\"stage_in\": %.0f, \"plan\": %.0f, \"request_to_collected\": %.0f}}\n",
                (unsigned long long)ctx->rt_steady_traced, (double)ctx->rt_steady_ns[0] / ctx->rt_steady_traced,
                (double)ctx->rt_steady_ns[1] / ctx->rt_steady_traced, (double)ctx->rt_steady_ns[2] / ctx->rt_steady_traced);
    cudaSetDevice(ctx->device);
    rt_quiesce(ctx);
    cudaDeviceSynchronize();
    if (ctx->rt_stream) cudaStreamDestroy(ctx->rt_stream);
    if (ctx->rt_mb) cudaFreeHost(ctx->rt_mb);
    for (Slot& sl : ctx->slots) {
        if (sl.d_in) cudaFree(sl.d_in);
        if (sl.d_out) cudaFree(sl.d_out);
        if (sl.h_in) cudaFreeHost(sl.h_in);
        if (sl.h_out) cudaFreeHost(sl.h_out);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    delete static_cast<CopyPool*>(ctx->copy_pool);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->done_flag) cudaFreeHost(ctx->done_flag);
    if (ctx->done_counter) cudaFree(ctx->done_counter);
    if (ctx->meta_stream) cudaStreamDestroy(ctx->meta_stream);
    for (MetaSlot& ms : ctx->meta) {
        if (ms.copied) cudaEventDestroy(ms.copied);
        if (ms.host) cudaFreeHost(ms.host);
        if (ms.dev) cudaFree(ms.dev);
        if (ms.done) cudaEventDestroy(ms.done);
    }
    if (ctx->tables_ready) cudaEventDestroy(ctx->tables_ready);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* doppler_b200_last_error(const doppler_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void* doppler_b200_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {   // portable: every device of a multi-GPU group DMAs from it
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void doppler_b200_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int doppler_b200_host_register(void* p, size_t bytes)
{
    if (!p || bytes == 0) return DOPPLER_B200_EINVAL;
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
        cudaGetLastError();
        return DOPPLER_B200_ECUDA;
    }
    return DOPPLER_B200_OK;
}

int doppler_b200_host_unregister(void* p)
{
    if (!p) return DOPPLER_B200_EINVAL;
    if (cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return DOPPLER_B200_ECUDA;
    }
    return DOPPLER_B200_OK;
}

uint64_t doppler_b200_launch_count(const doppler_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int doppler_b200_tune(doppler_b200_ctx* ctx, int knob, uint64_t value)
{
    if (!ctx) return DOPPLER_B200_EINVAL;
    switch (knob) {
    case DOPPLER_B200_TUNE_SMALL_MAX_SAMPLES:
        ctx->small_max = (uint32_t)std::min<uint64_t>(value, kLaunchMaxSamples);
        return DOPPLER_B200_OK;
    case DOPPLER_B200_TUNE_RESIDENT_IDLE_US:
        if (value > 10000000) return fail(ctx, DOPPLER_B200_EINVAL, "resident kernel idle time-out is limited to 10 s");
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        rt_quiesce(ctx);
        ctx->rt_idle_us = value;
        return DOPPLER_B200_OK;
    case DOPPLER_B200_TUNE_TINY_HOST_BYTES:
        if (value > (8u << 20)) return fail(ctx, DOPPLER_B200_EINVAL, "tiny host path is limited to 8 MiB");
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        rt_quiesce(ctx);
        abandon_slots(ctx);
        ctx->tiny_host_bytes = (size_t)value;
        return DOPPLER_B200_OK;
    case DOPPLER_B200_TUNE_SEG_VARIANT:
        ctx->seg_alt = value != 0;
        return DOPPLER_B200_OK;
    case DOPPLER_B200_TUNE_DECIM_VARIANT:
        ctx->decim_generic = value != 0;
        return DOPPLER_B200_OK;
    case DOPPLER_B200_TUNE_DECIM_STAGE_SLOTS:
        ctx->decim_stage_slots = (uint32_t)std::max<uint64_t>(512, std::min<uint64_t>(value, 27000));
        return DOPPLER_B200_OK;
    case DOPPLER_B200_TUNE_MAX_CLAIM:
        ctx->max_claim = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(value, 64));
        return DOPPLER_B200_OK;
    default:
        return fail(ctx, DOPPLER_B200_EINVAL, "unknown tuning knob %d", knob);
    }
}

int doppler_b200_synchronize(doppler_b200_ctx* ctx)
{
    if (!ctx) return DOPPLER_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DOPPLER_B200_OK;
}

// ---- fused path -------------------------------------------------------------------------------

int doppler_b200_mix_dev(doppler_b200_ctx* ctx, const void* d_in, size_t in_len, int intype, int outtype, float shift_hz,
                         uint32_t samplerate, uint32_t* samplenum, void* d_out, size_t out_cap, void* stream)
{
    uint64_t n = 0;
    int rc = check_common(ctx, d_in, in_len, intype, outtype, samplenum, d_out, out_cap, &n);
    if (rc) return rc;
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail(ctx, DOPPLER_B200_EINVAL, "device buffers must be 16-byte aligned");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<dplan::Run> runs{dplan::Run{n, dplan::ratio(shift_hz, samplerate)}};
    return launch_mix(ctx, d_in, d_out, n, intype, outtype, runs, samplenum, stream ? (cudaStream_t)stream : ctx->stream);
}

int doppler_b200_mix_blocks_dev(doppler_b200_ctx* ctx, const void* d_in, size_t in_len, int intype, int outtype,
                                const float* shifts, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                uint32_t* samplenum, void* d_out, size_t out_cap, void* stream)
{
    uint64_t n = 0, bs = 0;
    int rc = check_common(ctx, d_in, in_len, intype, outtype, samplenum, d_out, out_cap, &n);
    if (rc) return rc;
    if (n == 0) return DOPPLER_B200_OK;
    rc = check_blocks(ctx, in_len, intype, shifts, nblocks, block_bytes, &bs);
    if (rc) return rc;
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail(ctx, DOPPLER_B200_EINVAL, "device buffers must be 16-byte aligned");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<dplan::Run> runs = dplan::runs_from_blocks(shifts, nblocks, bs, samplerate, n);
    return launch_mix(ctx, d_in, d_out, n, intype, outtype, runs, samplenum, stream ? (cudaStream_t)stream : ctx->stream);
}

int doppler_b200_mix(doppler_b200_ctx* ctx, const void* in, size_t in_len, int intype, int outtype, float shift_hz,
                     uint32_t samplerate, uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len)
{
    uint64_t n = 0;
    int rc = check_common(ctx, in, in_len, intype, outtype, samplenum, out, out_cap, &n);
    if (rc) return rc;
    if (out_len) *out_len = 0;
    if (n == 0) return DOPPLER_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    rc = mix_host(ctx, in, n, intype, outtype, &shift_hz, 1, 0, samplerate, samplenum, out);
    if (rc == DOPPLER_B200_OK && out_len) *out_len = n * bytes_per_sample(outtype);
    return rc;
}

int doppler_b200_mix_blocks(doppler_b200_ctx* ctx, const void* in, size_t in_len, int intype, int outtype,
                            const float* shifts, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                            uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len)
{
    uint64_t n = 0, bs = 0;
    int rc = check_common(ctx, in, in_len, intype, outtype, samplenum, out, out_cap, &n);
    if (rc) return rc;
    if (out_len) *out_len = 0;
    if (n == 0) return DOPPLER_B200_OK;
    rc = check_blocks(ctx, in_len, intype, shifts, nblocks, block_bytes, &bs);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    rc = mix_host(ctx, in, n, intype, outtype, shifts, nblocks, bs, samplerate, samplenum, out);
    if (rc == DOPPLER_B200_OK && out_len) *out_len = n * bytes_per_sample(outtype);
    return rc;
}

int doppler_b200_pipeline_probe(doppler_b200_ctx* ctx, const void* in, size_t in_len, int intype, int outtype, void* out, size_t out_cap)
{
    uint64_t n = 0;
    uint32_t sn = 0;
    int rc = check_common(ctx, in, in_len, intype, outtype, &sn, out, out_cap, &n);
    if (rc || n == 0) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const float zero = 0.0f;
    return mix_host(ctx, in, n, intype, outtype, &zero, 1, 0, 1, &sn, out, /*copy_only=*/true);
}

// ---- fused downstream stage: mix + decimating FIR ---------------------------
}  // extern "C"

struct doppler_b200_decim {
    doppler_b200_ctx* ctx = nullptr;
    uint32_t ntaps = 0, M = 0;
    float* d_taps = nullptr;
    float2* d_hist[2] = {nullptr, nullptr};   // history ping-pong (ntaps-1 mixed samples)
    int cur = 0;
    uint64_t pos = 0;                          // stream position of the next input sample
    DevPiece* d_pieces = nullptr;              // piece list of the call in flight (when it does not fit the kernel parameters)
    size_t pieces_cap = 0;
    cudaEvent_t hist_ready = nullptr;          // the history of the latest call has been written
    cudaStream_t hist_stream = nullptr;
    bool hist_event_valid = false;
    // register-blocked kernel (decimate_kernels.cuh): tap layout and walk segments of this filter, or fast_ok == false
    bool fast_ok = false;
    struct Scratch {                    // two-pass form: the mixed stream of one call (history in front), one buffer per stream in use
        cudaStream_t stream;
        float2* buf;
        size_t cap;                     // samples
    };
    std::vector<Scratch> scratch;
    uint32_t cuts[8] = {};              // sorted bounds of the walk's segments
    dmix::DecimFastArgs fast;
};

namespace {

constexpr uint32_t kDecimMaxTaps = 4096;
constexpr uint32_t kDecimStageSlots = 5120;   // mixed samples staged per CTA step (40 KB)

using DecimKernel = void (*)(const dmix::DecimArgs);
using dmix::DecimFastKernel;

// Register-blocked kernel (decimate_kernels.cuh): what depends on the filter alone -- the tap layout per walk position, the
// sorted segment bounds, the shape.  The envelope: the layout fits the kernel parameters.
void decim_fast_setup(doppler_b200_decim* d, const float* taps)
{
    constexpr uint32_t R = dmix::kDfR;
    const uint32_t M = d->M, ntaps = d->ntaps, ntq = (R - 1) * M + ntaps;
    d->fast_ok = false;
    if (M > 64 || ntq > (uint32_t)dmix::kDfMaxTq) return;
    dmix::DecimFastArgs& f = d->fast;
    memset(&f, 0, sizeof f);
    for (uint32_t u = 0; u < ntq; u++)
        for (uint32_t k = 0; k < R; k++) {
            const int64_t t = (int64_t)u - (int64_t)(R - 1 - k) * M;
            const float h = (t >= 0 && t < (int64_t)ntaps) ? taps[t] : 0.0f;
            uint32_t b;
            memcpy(&b, &h, 4);
            f.tq[u].h[k] = (uint64_t)b | ((uint64_t)b << 32);
        }
    // output k is active at walk positions [(3 - k) * M, (3 - k) * M + ntaps): the sorted bounds cut the walk into 7 segments
    for (uint32_t k = 0; k < R; k++) {
        d->cuts[k] = k * M;
        d->cuts[R + k] = k * M + ntaps;
    }
    std::sort(d->cuts, d->cuts + 2 * R);
    f.shape = std::min<uint32_t>(3, (ntaps - 1) / M);
    f.rm_magic = (uint32_t)((1ull << 32) / (R * M)) + 1u;
    d->fast_ok = true;
}

// ... and what depends on the call: threads per CTA from the stage budget, the staging origin, and with it every walk position's
// slot offset.  False when the launch does not fit (the generic kernel takes it).
bool decim_fast_plan(doppler_b200_decim* d, uint64_t i0, uint32_t stage_slots, const std::vector<DevPiece>& pieces, size_t* smem_bytes,
                     bool mixed = false)
{
    constexpr uint32_t R = dmix::kDfR;
    const uint32_t M = d->M, ntaps = d->ntaps, RM = R * M;
    dmix::DecimFastArgs& f = d->fast;
    f.lead = (uint32_t)((((int64_t)i0 - (int64_t)(ntaps - 1)) % 4 + 4) % 4);
    // a CTA step stages lead + (4 * tb - 1) * M + ntaps samples (+ one 16-byte group of slack), one padding slot per 4M
    uint32_t tb = dmix::kDfMaxThreads;
    uint64_t slots = 0;
    for (; tb >= 32; tb -= 32) {
        const uint64_t count = f.lead + (uint64_t)(R * tb - 1) * M + ntaps + 4;
        slots = count + count / RM + 2;
        if (slots <= stage_slots) break;
    }
    if (tb < 32) return false;
    f.tb = tb;
    // walk position u reads staged index c0 - u of the thread's group: its slot offset (one padding slot per 4M) rides with its taps
    const uint32_t c0 = f.lead + (ntaps - 1) + (R - 1) * M;
    for (int i = 0; i < 8; i++) f.cuts[i] = d->cuts[i];
    for (uint32_t u = 0; u < (R - 1) * M + ntaps; u++) {
        const uint32_t c = c0 - u;
        f.tq[u].off = (c + c / RM) * 8u;
    }
    // the longest tabled period that fits the CTA's table area decides its size; a launch whose samples mostly lie in pieces
    // WITHOUT a table (long periods, track mode) gains nothing from this kernel's staging loop and keeps the generic kernel
    uint32_t cap = 0;
    uint64_t tabled = 0, total = 0;
    for (const DevPiece& p : pieces) {
        total += p.k_end - p.k_begin;
        if (p.tab == dmix::kNoTab) continue;
        tabled += p.k_end - p.k_begin;
        if (p.period + dmix::kTabPad <= dmix::kDfTabCap) cap = std::max<uint32_t>(cap, p.period + dmix::kTabPad);
    }
    if (mixed) {
        cap = 0;   // (input that is already mixed: no pieces, no table)
    } else if (2 * tabled < total) {
        return false;
    }
    f.tab_cap = (cap + 1) & ~1u;
    *smem_bytes = (size_t)f.tab_cap * sizeof(float2) + (size_t)slots * sizeof(float2);
    return true;
}

// The two-pass form's scratch buffer for launches on stream `s`: at least `samples` complex f32.
float2* decim_scratch(doppler_b200_decim* dec, cudaStream_t s, size_t samples)
{
    for (auto& sc : dec->scratch)
        if (sc.stream == s) {
            if (sc.cap >= samples) return sc.buf;
            rt_quiesce(dec->ctx);
            cudaStreamSynchronize(s);
            cudaFree(sc.buf);
            sc.buf = nullptr;
            sc.cap = 0;
            if (cudaMalloc(&sc.buf, samples * sizeof(float2)) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            sc.cap = samples;
            return sc.buf;
        }
    float2* buf = nullptr;
    if (cudaMalloc(&buf, samples * sizeof(float2)) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    dec->scratch.push_back({s, buf, samples});
    return buf;
}

// One device-resident call of the fused stage: `runs` over n samples at d_in; outputs to d_out.  Asynchronous on `s`.
int decimate_launch(doppler_b200_decim* dec, const void* d_in, uint64_t n, int intype, int outtype, const std::vector<dplan::Run>& runs,
                    uint32_t* samplenum, void* d_out, uint64_t* nout_ret, cudaStream_t s)
{
    doppler_b200_ctx* ctx = dec->ctx;
    *nout_ret = 0;
    if (n == 0) return DOPPLER_B200_OK;
    if (n > kLaunchMaxSamples) return fail(ctx, DOPPLER_B200_EINVAL, "one fused mix + decimate call is limited to 2^30 samples");
    std::vector<dplan::Piece> pieces;
    uint32_t sn_after = *samplenum;
    ctx->planner.plan(runs, 0, &sn_after, &pieces);
    std::vector<DevPiece> dev;
    std::vector<const dplan::Piece*> src;
    clip_pieces(pieces, 0, n, 1, &dev, &src);
    if (ctx->tables_event_valid && s != ctx->tables_stream) CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->tables_ready, 0));
    for (size_t i = 0; i < dev.size(); i++)
        if (dev[i].period) {
            int rc = get_table(ctx, dev[i].r, dev[i].period, src[i]->k_end - src[i]->k_begin, s, &dev[i].tab);
            if (rc) return rc;
        }
    for (DevPiece& d : dev) {   // a recycled arena invalidates offsets taken earlier in this call
        if (d.tab == dmix::kNoTab) continue;
        uint32_t key;
        memcpy(&key, &d.r, 4);
        auto it = ctx->tables.find(key);
        d.tab = (it != ctx->tables.end() && it->second.period == d.period) ? it->second.off : dmix::kNoTab;
    }
    dmix::DecimArgs a;
    memset(&a, 0, sizeof a);
    a.mix.in = d_in;
    a.mix.tables = ctx->arena;
    a.mix.nsamples = (uint32_t)n;
    a.mix.npieces = (uint32_t)dev.size();
    if (dev.size() <= (size_t)dmix::kInlinePieces) {
        for (size_t i = 0; i < dev.size(); i++) a.mix.inl[i] = dev[i];
    } else {
        if (dec->pieces_cap < dev.size()) {
            if (dec->d_pieces) CUDA_TRY(ctx, cudaFree(dec->d_pieces));
            dec->d_pieces = nullptr;
            dec->pieces_cap = 0;
            CUDA_TRY(ctx, cudaMalloc(&dec->d_pieces, dev.size() * 2 * sizeof(DevPiece)));
            dec->pieces_cap = dev.size() * 2;
        }
        // pageable source: the copy is staged before the call returns, so `dev` may go out of scope
        CUDA_TRY(ctx, cudaMemcpyAsync(dec->d_pieces, dev.data(), dev.size() * sizeof(DevPiece), cudaMemcpyHostToDevice, s));
        a.mix.pieces = dec->d_pieces;
    }
    // the previous call's history must have landed (it may have been written on another slot's stream)
    if (dec->hist_event_valid && s != dec->hist_stream) CUDA_TRY(ctx, cudaStreamWaitEvent(s, dec->hist_ready, 0));
    const uint32_t M = dec->M;
    a.out = d_out;
    a.hist = dec->d_hist[dec->cur];
    a.hist_next = dec->d_hist[dec->cur ^ 1];
    a.taps = dec->d_taps;
    a.ntaps = dec->ntaps;
    a.M = M;
    const uint64_t i0 = (M - dec->pos % M) % M;
    a.first_out = (uint32_t)i0;
    const uint64_t nout = i0 < n ? (n - i0 + M - 1) / M : 0;
    a.nout = (uint32_t)nout;
    a.skew = (M % 2 == 0) ? 1u : 0u;
    // outputs per CTA step: as many as the staging area holds, at most one per thread
    const uint64_t usable = a.skew ? (uint64_t)kDecimStageSlots * M / (M + 1) - 1 : kDecimStageSlots;
    uint64_t ot = usable > dec->ntaps ? (usable - dec->ntaps) / M + 1 : 1;
    ot = std::max<uint64_t>(1, std::min<uint64_t>(ot, dmix::kDecimThreads));
    a.out_per_cta = (uint32_t)ot;
    const uint64_t count = (ot - 1) * M + dec->ntaps;
    const size_t slots = a.skew ? count + count / M + 2 : count;
    const size_t smem = ((dec->ntaps * 4 + 15) & ~(size_t)15) + slots * sizeof(float2);
    static const DecimKernel kern[2][2] = {{dmix::mix_decimate_kernel<0, 0>, dmix::mix_decimate_kernel<0, 1>},
                                           {dmix::mix_decimate_kernel<1, 0>, dmix::mix_decimate_kernel<1, 1>}};
    size_t fsmem = 0;
    const bool can_fast = nout && dec->fast_ok && !ctx->decim_generic && ((((uintptr_t)d_in) | ((uintptr_t)d_out)) & 15) == 0;
    auto launch_fast = [&](int in_kind) -> int {   // in_kind: 0 i16, 1 f32, 2 already mixed
        constexpr uint32_t R = dmix::kDfR;
        dmix::DecimFastArgs& f = dec->fast;
        const uint64_t steps = (nout + (uint64_t)R * f.tb - 1) / ((uint64_t)R * f.tb);
        const uint32_t nt = f.tb > 128 ? 256 : 128;   // CTA size: the next instantiated size that holds the output-owning threads
        const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(2048 / nt, (size_t)(227 * 1024) / (fsmem + 1024)));
        const uint32_t grid = (uint32_t)std::min<uint64_t>(steps, (uint64_t)ctx->sm_count * per_sm);
        static DecimFastKernel (*const pick[3][2])(int, int) = {{dmix::df_kernel_0_0, dmix::df_kernel_0_1},
                                                                {dmix::df_kernel_1_0, dmix::df_kernel_1_1},
                                                                {dmix::df_kernel_2_0, dmix::df_kernel_2_1}};
        const DecimFastKernel fk = pick[in_kind][outtype == DOPPLER_B200_I16 ? 0 : 1](nt == 128, (int)f.shape);
        CUDA_TRY(ctx, cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        fk<<<grid, nt, fsmem, s>>>(f);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
        return DOPPLER_B200_OK;
    };
    bool history_done = false;
    constexpr uint64_t kTwoPassMax = 1ull << 28;   // 2 GiB of scratch at most
    if (can_fast && decim_fast_plan(dec, i0, ctx->decim_stage_slots, dev, &fsmem)) {
        // register-blocked kernel: 4 consecutive outputs per thread, taps in the kernel parameters (decimate_kernels.cuh)
        dec->fast.d = a;
        int rc = launch_fast(intype);
        if (rc) return rc;
    } else if (can_fast && n <= kTwoPassMax && decim_fast_plan(dec, i0 + ((dec->ntaps - 1 + 3) & ~3u), ctx->decim_stage_slots, dev, &fsmem, /*mixed=*/true)) {
        // Mostly table-less pieces (long periods, track mode): the mixer itself writes the mixed stream as complex f32 into a
        // scratch buffer, behind the carried history, and the register-blocked kernel runs over that (decimate_kernels.cuh).
        const uint32_t nh = dec->ntaps - 1, H = (nh + 3) & ~3u;   // the history ends on a 32-byte boundary
        float2* y = decim_scratch(dec, s, (size_t)H + n);
        if (!y) return fail(ctx, DOPPLER_B200_ECUDA, "decimator: no memory for %llu samples of scratch", (unsigned long long)(H + n));
        if (nh) CUDA_TRY(ctx, cudaMemcpyAsync(y + (H - nh), dec->d_hist[dec->cur], nh * sizeof(float2), cudaMemcpyDeviceToDevice, s));
        uint32_t sn_mix = *samplenum;
        int rc = launch_mix(ctx, d_in, y + H, n, intype, DOPPLER_B200_F32, runs, &sn_mix, s);
        if (rc) return rc;
        dmix::DecimArgs b = a;
        memset(&b.mix, 0, sizeof b.mix);
        b.mix.in = y;
        b.mix.nsamples = (uint32_t)(H + n);
        b.first_out = (uint32_t)(i0 + H);
        dec->fast.d = b;
        rc = launch_fast(2);
        if (rc) return rc;
        // the next call's history: the last ntaps-1 mixed samples (older ones from this call's history when n is short)
        if (nh) CUDA_TRY(ctx, cudaMemcpyAsync(dec->d_hist[dec->cur ^ 1], y + (H + n - nh), nh * sizeof(float2), cudaMemcpyDeviceToDevice, s));
        history_done = true;
    } else if (nout) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>((nout + ot - 1) / ot, (uint64_t)ctx->sm_count * 8);
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern[intype][outtype], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern[intype][outtype]<<<grid, dmix::kDecimThreads, smem, s>>>(a);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    if (dec->ntaps > 1) {
        if (!history_done) {
            const uint32_t hgrid = (dec->ntaps - 1 + dmix::kDecimThreads - 1) / dmix::kDecimThreads;
            if (intype == DOPPLER_B200_I16)
                dmix::decim_history_kernel<0><<<hgrid, dmix::kDecimThreads, 0, s>>>(a);
            else
                dmix::decim_history_kernel<1><<<hgrid, dmix::kDecimThreads, 0, s>>>(a);
            CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches++;
        }
        CUDA_TRY(ctx, cudaEventRecord(dec->hist_ready, s));
        dec->hist_event_valid = true;
        dec->hist_stream = s;
        dec->cur ^= 1;
    }
    dec->pos += n;
    *samplenum = sn_after;
    *nout_ret = nout;
    return DOPPLER_B200_OK;
}

int decim_check(doppler_b200_decim* dec, const void* in, size_t in_len, int intype, int outtype, uint32_t* samplenum, void* out, size_t out_cap,
                uint64_t* n, uint64_t* nout)
{
    if (!dec) return DOPPLER_B200_EINVAL;
    doppler_b200_ctx* ctx = dec->ctx;
    if (!valid_type(intype) || !valid_type(outtype)) return fail(ctx, DOPPLER_B200_EINVAL, "unknown IQ data type");
    if (!samplenum) return fail(ctx, DOPPLER_B200_EINVAL, "samplenum is NULL");
    const size_t ibps = bytes_per_sample(intype);
    if (in_len % ibps != 0) return fail(ctx, DOPPLER_B200_EALIGN, "input length %zu is not a multiple of %zu (dsp.rs assert)", in_len, ibps);
    *n = in_len / ibps;
    const uint64_t i0 = (dec->M - dec->pos % dec->M) % dec->M;
    *nout = i0 < *n ? (*n - i0 + dec->M - 1) / dec->M : 0;
    if (*n && !in) return fail(ctx, DOPPLER_B200_EINVAL, "NULL buffer");
    if (*nout && !out) return fail(ctx, DOPPLER_B200_EINVAL, "NULL buffer");
    if (*nout * bytes_per_sample(outtype) > out_cap)
        return fail(ctx, DOPPLER_B200_ECAP, "output capacity %zu < %llu bytes needed", out_cap, (unsigned long long)(*nout * bytes_per_sample(outtype)));
    return DOPPLER_B200_OK;
}

std::vector<dplan::Run> decim_runs(const float* shifts, size_t nblocks, uint64_t block_samples, uint32_t samplerate, uint64_t k, uint64_t n)
{
    if (nblocks <= 1 || block_samples == 0) return {dplan::Run{n, dplan::ratio(shifts[0], samplerate)}};
    const size_t b0 = (size_t)(k / block_samples);
    return dplan::runs_from_blocks(shifts + b0, nblocks - b0, block_samples, samplerate, n);
}

// Host buffers: the chunk pipeline of mix_host (H2D / kernels / D2H of neighbouring chunks overlap across the slots' streams;
// the history hand-over between chunks is ordered through dec->hist_ready).
int decimate_host(doppler_b200_decim* dec, const void* in, uint64_t n, int intype, int outtype, const float* shifts, size_t nblocks,
                  uint64_t block_samples, uint32_t samplerate, uint32_t* samplenum, void* out, size_t* out_len)
{
    doppler_b200_ctx* ctx = dec->ctx;
    const size_t ibps = bytes_per_sample(intype), obps = bytes_per_sample(outtype);
    uint64_t chunk = kHostChunkBytes / ibps;
    if (block_samples && nblocks > 1) chunk = std::max<uint64_t>(block_samples, chunk / block_samples * block_samples);
    const bool in_pinned = is_pinned(in), out_pinned = is_pinned(out);
    uint32_t sn = *samplenum;
    size_t written = 0;
    int c = 0;
    for (uint64_t k = 0; k < n; k += chunk, c++) {
        const uint64_t m = std::min(chunk, n - k);
        Slot& sl = ctx->slots[c % kSlots];
        int rc = retire_slot(ctx, sl);
        if (rc) return rc;
        rc = ensure_slot(ctx, sl, std::min<uint64_t>(chunk, n) * ibps, (std::min<uint64_t>(chunk, n) / dec->M + 2) * obps);
        if (rc) return rc;
        const char* src = static_cast<const char*>(in) + k * ibps;
        if (!in_pinned) {
            staged_copy(ctx, sl.h_in, src, m * ibps);
            src = static_cast<const char*>(sl.h_in);
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(sl.d_in, src, m * ibps, cudaMemcpyHostToDevice, sl.stream));
        uint64_t nout = 0;
        rc = decimate_launch(dec, sl.d_in, m, intype, outtype, decim_runs(shifts, nblocks, block_samples, samplerate, k, m), &sn, sl.d_out, &nout,
                             sl.stream);
        if (rc) return rc;
        char* dst = static_cast<char*>(out) + written;
        if (nout) {
            if (out_pinned) {
                CUDA_TRY(ctx, cudaMemcpyAsync(dst, sl.d_out, nout * obps, cudaMemcpyDeviceToHost, sl.stream));
                sl.user_out = nullptr;
            } else {
                CUDA_TRY(ctx, cudaMemcpyAsync(sl.h_out, sl.d_out, nout * obps, cudaMemcpyDeviceToHost, sl.stream));
                sl.user_out = dst;
                sl.user_out_bytes = nout * obps;
            }
        }
        CUDA_TRY(ctx, cudaEventRecord(sl.done, sl.stream));
        sl.busy = true;
        written += nout * obps;
    }
    for (int i = 0; i < kSlots; i++) {
        int rc = retire_slot(ctx, ctx->slots[i]);
        if (rc) return rc;
    }
    *samplenum = sn;
    if (out_len) *out_len = written;
    return DOPPLER_B200_OK;
}

}  // namespace

extern "C" {

int doppler_b200_decim_create(doppler_b200_ctx* ctx, const float* taps, uint32_t ntaps, uint32_t decimation, doppler_b200_decim** out)
{
    if (!ctx || !out) return DOPPLER_B200_EINVAL;
    *out = nullptr;
    if (!taps || ntaps == 0 || ntaps > kDecimMaxTaps || decimation == 0 || decimation > (1u << 20))
        return fail(ctx, DOPPLER_B200_EINVAL, "decimator: 1..%u taps, decimation 1..2^20", kDecimMaxTaps);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    doppler_b200_decim* d = new (std::nothrow) doppler_b200_decim;
    if (!d) return DOPPLER_B200_ENOMEM;
    d->ctx = ctx;
    d->ntaps = ntaps;
    d->M = decimation;
    const size_t hb = std::max<size_t>(ntaps - 1, 1) * sizeof(float2);
    cudaError_t e = cudaMalloc(&d->d_taps, ntaps * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d->d_taps, taps, ntaps * sizeof(float), cudaMemcpyHostToDevice);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaMalloc(&d->d_hist[i], hb);
        if (e == cudaSuccess) e = cudaMemset(d->d_hist[i], 0, hb);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->hist_ready, cudaEventDisableTiming);
    decim_fast_setup(d, taps);
    if (e != cudaSuccess) {
        fail(ctx, DOPPLER_B200_ECUDA, "decimator setup failed: %s", cudaGetErrorString(e));
        doppler_b200_decim_destroy(d);
        return DOPPLER_B200_ECUDA;
    }
    *out = d;
    return DOPPLER_B200_OK;
}

void doppler_b200_decim_destroy(doppler_b200_decim* d)
{
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    rt_quiesce(d->ctx);
    cudaDeviceSynchronize();
    if (d->d_taps) cudaFree(d->d_taps);
    for (float2* h : d->d_hist)
        if (h) cudaFree(h);
    if (d->d_pieces) cudaFree(d->d_pieces);
    for (auto& sc : d->scratch)
        if (sc.buf) cudaFree(sc.buf);
    if (d->hist_ready) cudaEventDestroy(d->hist_ready);
    delete d;
}

int doppler_b200_decim_reset(doppler_b200_decim* d)
{
    if (!d) return DOPPLER_B200_EINVAL;
    CUDA_TRY(d->ctx, cudaSetDevice(d->ctx->device));
    rt_quiesce(d->ctx);
    CUDA_TRY(d->ctx, cudaDeviceSynchronize());
    const size_t hb = std::max<size_t>(d->ntaps - 1, 1) * sizeof(float2);
    for (float2* h : d->d_hist) CUDA_TRY(d->ctx, cudaMemset(h, 0, hb));
    d->pos = 0;
    d->cur = 0;
    d->hist_event_valid = false;
    return DOPPLER_B200_OK;
}

uint64_t doppler_b200_decim_position(const doppler_b200_decim* d) { return d ? d->pos : 0; }

long doppler_b200_decim_walk_trace(const float* taps, uint32_t ntaps, uint32_t decimation, uint64_t first_out, uint32_t* records,
                                   uint32_t* tap_bits, size_t cap, uint32_t* info)
{
    if (!taps || ntaps == 0 || ntaps > kDecimMaxTaps || decimation == 0 || !records) return -1;
    std::unique_ptr<doppler_b200_decim> d(new (std::nothrow) doppler_b200_decim);
    if (!d) return -1;
    d->ntaps = ntaps;
    d->M = decimation;
    decim_fast_setup(d.get(), taps);
    size_t smem = 0;
    uint32_t nt = 0;
    // (a launch of tabled pieces only: one piece of period 100 covering the call)
    std::vector<DevPiece> pieces(1);
    memset(&pieces[0], 0, sizeof(DevPiece));
    pieces[0].k_end = 1u << 20;
    pieces[0].period = 100;
    pieces[0].tab = 0;
    if (!d->fast_ok || !decim_fast_plan(d.get(), first_out % decimation, dmix::kDfStageSlots, pieces, &smem)) return 0;
    nt = d->fast.tb > 128 ? 256 : 128;
    const dmix::DecimFastArgs& f = d->fast;
    constexpr uint32_t R = dmix::kDfR;
    const uint32_t M = decimation, RM = R * M;
    static const int lo[4][7] = {{3, 1, 2, 1, 1, 1, 0}, {3, 2, 2, 1, 1, 0, 0}, {3, 2, 1, 1, 0, 0, 0}, {3, 2, 1, 0, 0, 0, 0}};
    static const int hi[4][7] = {{3, 0, 2, 0, 1, 0, 0}, {3, 3, 2, 2, 1, 1, 0}, {3, 3, 3, 2, 2, 1, 0}, {3, 3, 3, 3, 2, 1, 0}};
    // the device's ranges (decimate_kernels.cuh: DfRange) are the source of truth: the tables above must agree with them
    static_assert(dmix::DfRange<3, 3>::lo == 0 && dmix::DfRange<3, 3>::hi == 3 && dmix::DfRange<0, 1>::lo == 1 && dmix::DfRange<0, 1>::hi == 0 &&
                      dmix::DfRange<1, 2>::lo == 2 && dmix::DfRange<1, 2>::hi == 2 && dmix::DfRange<2, 3>::lo == 1 && dmix::DfRange<2, 3>::hi == 2,
                  "walk ranges");
    const uint32_t c0 = f.lead + (ntaps - 1) + (R - 1) * M;
    size_t n = 0;
    for (int i = 0; i < 7; i++) {
        for (uint32_t u = f.cuts[i]; u < f.cuts[i + 1]; u++, n++) {
            if (n >= cap) continue;
            uint32_t* rec = records + n * 8;
            rec[0] = c0 - u;
            rec[1] = f.tq[u].off / 8u;   // relative to the thread's base slot tid * (4M + 1)
            rec[2] = (uint32_t)lo[f.shape][i];
            rec[3] = (uint32_t)hi[f.shape][i];
            for (uint32_t k = 0; k < R; k++) {
                const bool on = (int)k >= lo[f.shape][i] && (int)k <= hi[f.shape][i];
                rec[4 + k] = on ? u - (R - 1 - k) * M : 0xffffffffu;
                if (tap_bits) tap_bits[n * 4 + k] = (uint32_t)f.tq[u].h[k];
            }
        }
    }
    (void)RM;
    if (info) {
        info[0] = f.tb;
        info[1] = f.lead;
        info[2] = nt;
        info[3] = f.shape;
    }
    return (long)n;
}

int doppler_b200_mix_blocks_decimate(doppler_b200_decim* dec, const void* in, size_t in_len, int intype, int outtype,
                                     const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                     uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len)
{
    uint64_t n = 0, nout = 0, bs = 0;
    int rc = decim_check(dec, in, in_len, intype, outtype, samplenum, out, out_cap, &n, &nout);
    if (rc) return rc;
    if (out_len) *out_len = 0;
    if (n == 0) return DOPPLER_B200_OK;
    rc = check_blocks(dec->ctx, in_len, intype, shift_hz_per_block, nblocks, block_bytes, &bs);
    if (rc) return rc;
    CUDA_TRY(dec->ctx, cudaSetDevice(dec->ctx->device));
    const uint64_t pos0 = dec->pos;
    const int cur0 = dec->cur;
    rc = decimate_host(dec, in, n, intype, outtype, shift_hz_per_block, nblocks, bs, samplerate, samplenum, out, out_len);
    if (rc) {   // nothing of a failed call may stay in flight or in the state
        const std::string keep = dec->ctx->err;
        abandon_slots(dec->ctx);
        dec->ctx->err = keep;
        dec->pos = pos0;
        dec->cur = cur0;
    }
    return rc;
}

int doppler_b200_mix_decimate(doppler_b200_decim* dec, const void* in, size_t in_len, int intype, int outtype, float shift_hz,
                              uint32_t samplerate, uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len)
{
    const size_t whole = in_len ? in_len : 1;
    return doppler_b200_mix_blocks_decimate(dec, in, in_len, intype, outtype, &shift_hz, 1, (whole + 7) / 8 * 8, samplerate, samplenum, out, out_cap,
                                            out_len);
}

int doppler_b200_mix_blocks_decimate_dev(doppler_b200_decim* dec, const void* d_in, size_t in_len, int intype, int outtype,
                                         const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                         uint32_t* samplenum, void* d_out, size_t out_cap, size_t* out_len, void* stream)
{
    uint64_t n = 0, nout = 0, bs = 0;
    int rc = decim_check(dec, d_in, in_len, intype, outtype, samplenum, d_out, out_cap, &n, &nout);
    if (rc) return rc;
    if (out_len) *out_len = 0;
    if (n == 0) return DOPPLER_B200_OK;
    rc = check_blocks(dec->ctx, in_len, intype, shift_hz_per_block, nblocks, block_bytes, &bs);
    if (rc) return rc;
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail(dec->ctx, DOPPLER_B200_EINVAL, "device buffers must be 16-byte aligned");
    CUDA_TRY(dec->ctx, cudaSetDevice(dec->ctx->device));
    uint64_t got = 0;
    rc = decimate_launch(dec, d_in, n, intype, outtype, decim_runs(shift_hz_per_block, nblocks, bs, samplerate, 0, n), samplenum, d_out, &got,
                         stream ? (cudaStream_t)stream : dec->ctx->stream);
    if (rc == DOPPLER_B200_OK && out_len) *out_len = got * bytes_per_sample(outtype);
    return rc;
}

int doppler_b200_mix_decimate_dev(doppler_b200_decim* dec, const void* d_in, size_t in_len, int intype, int outtype, float shift_hz,
                                  uint32_t samplerate, uint32_t* samplenum, void* d_out, size_t out_cap, size_t* out_len, void* stream)
{
    const size_t whole = in_len ? in_len : 1;
    return doppler_b200_mix_blocks_decimate_dev(dec, d_in, in_len, intype, outtype, &shift_hz, 1, (whole + 7) / 8 * 8, samplerate, samplenum, d_out,
                                                out_cap, out_len, stream);
}

// ---- the reference's three functions, one to one ---------------------------------------------

int doppler_b200_shift_frequency(doppler_b200_ctx* ctx, const float* inbuf, size_t nsamples, uint32_t* samplenum,
                                 float shift_hz, uint32_t samplerate, float* out)
{
    // Complex<f32> in, Complex<f32> out: the f32 ingest is a bit copy (dsp.rs:101-115) and the f32
    // egress is a byte view (main.rs:89-93), so this IS the fused path with both types f32.
    return doppler_b200_mix(ctx, inbuf, nsamples * 8, DOPPLER_B200_F32, DOPPLER_B200_F32, shift_hz, samplerate, samplenum,
                            out, nsamples * 8, nullptr);
}

static int convert_host(doppler_b200_ctx* ctx, const uint8_t* inbuf, size_t len, int intype, float* out)
{
    if (!ctx) return DOPPLER_B200_EINVAL;
    const size_t ibps = bytes_per_sample(intype);
    if (len % ibps != 0)
        return fail(ctx, DOPPLER_B200_EALIGN, "input length %zu is not a multiple of %zu (dsp.rs assert)", len, ibps);
    const uint64_t n = len / ibps;
    if (n == 0) return DOPPLER_B200_OK;
    if (!inbuf || !out) return fail(ctx, DOPPLER_B200_EINVAL, "NULL buffer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t chunk = kHostChunkBytes / ibps;
    for (uint64_t k = 0; k < n; k += chunk) {
        const uint64_t m = std::min(chunk, n - k);
        Slot& sl = ctx->slots[0];
        int rc = retire_slot(ctx, sl);   // a pending copy-out of an earlier call must not outlive the buffers ensure_slot may replace
        if (rc == DOPPLER_B200_OK) rc = ensure_slot(ctx, sl, std::min(chunk, n) * ibps, std::min(chunk, n) * 8);
        if (rc) {
            abandon_slots(ctx);
            return rc;
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(sl.d_in, inbuf + k * ibps, m * ibps, cudaMemcpyHostToDevice, sl.stream));
        const uint32_t grid = (uint32_t)std::min<uint64_t>((m + dmix::kThreads - 1) / dmix::kThreads, (uint64_t)ctx->sm_count * 32);
        if (intype == DOPPLER_B200_I16)
            dmix::convert_kernel<0><<<grid, dmix::kThreads, 0, sl.stream>>>(sl.d_in, (float2*)sl.d_out, (uint32_t)m);
        else
            dmix::convert_kernel<1><<<grid, dmix::kThreads, 0, sl.stream>>>(sl.d_in, (float2*)sl.d_out, (uint32_t)m);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
        CUDA_TRY(ctx, cudaMemcpyAsync(out + 2 * k, sl.d_out, m * 8, cudaMemcpyDeviceToHost, sl.stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(sl.stream));
    }
    return DOPPLER_B200_OK;
}

int doppler_b200_convert_iqi16_to_complex(doppler_b200_ctx* ctx, const uint8_t* inbuf, size_t len, float* out)
{
    return convert_host(ctx, inbuf, len, DOPPLER_B200_I16, out);
}

int doppler_b200_convert_iqf32_to_complex(doppler_b200_ctx* ctx, const uint8_t* inbuf, size_t len, float* out)
{
    return convert_host(ctx, inbuf, len, DOPPLER_B200_F32, out);
}

// ---- analytic samplenum (host only) -----------------------------------------------------------

uint32_t doppler_b200_samplenum_advance(uint32_t samplenum, float shift_hz, uint32_t samplerate, uint64_t count)
{
    dplan::Planner pl;
    std::vector<dplan::Run> runs{dplan::Run{count, dplan::ratio(shift_hz, samplerate)}};
    return pl.advance(runs, samplenum);
}

uint32_t doppler_b200_samplenum_advance_blocks(uint32_t samplenum, const float* shifts, size_t nblocks, uint64_t block_samples,
                                               uint32_t samplerate, uint64_t count)
{
    if (!shifts || nblocks == 0 || block_samples == 0) return samplenum;
    dplan::Planner pl;
    return pl.advance(dplan::runs_from_blocks(shifts, nblocks, block_samples, samplerate, count), samplenum);
}

long doppler_b200_plan_trace(uint32_t* samplenum, const float* shifts, size_t nblocks, uint64_t block_samples,
                             uint32_t samplerate, uint64_t count, uint32_t* trace)
{
    if (!samplenum || !shifts || nblocks == 0 || block_samples == 0) return -1;
    dplan::Planner pl;
    std::vector<dplan::Piece> pieces;
    pl.plan(dplan::runs_from_blocks(shifts, nblocks, block_samples, samplerate, count), 0, samplenum, &pieces);
    if (trace) {
        for (const dplan::Piece& p : pieces)
            for (uint64_t k = p.k_begin; k < p.k_end; k++) {
                const uint64_t off = k - p.k_begin;
                trace[k] = p.period ? (uint32_t)(((uint64_t)p.base + off) % p.period) + 1u : p.base + (uint32_t)off;
            }
    }
    return (long)pieces.size();
}

long doppler_b200_plan_tiles_trace(int intype, int outtype, uint32_t samplenum, const float* shifts, size_t nblocks,
                                   uint64_t block_samples, uint32_t samplerate, uint64_t count, uint32_t npipes,
                                   uint32_t* trace, uint32_t* cover, uint64_t* stats)
{
    if (!valid_type(intype) || !valid_type(outtype) || !shifts || nblocks == 0 || block_samples == 0 || !trace || !cover ||
        npipes == 0 || count == 0 || count > kLaunchMaxSamples)
        return -1;
    dplan::Planner pl;
    std::vector<dplan::Piece> pieces;
    pl.plan(dplan::runs_from_blocks(shifts, nblocks, block_samples, samplerate, count), 0, &samplenum, &pieces);
    std::vector<DevPiece> dev;
    clip_pieces(pieces, 0, count, shape_for(intype, outtype).seg.row_samples, &dev, nullptr);
    if (intype == DOPPLER_B200_I16 && outtype == DOPPLER_B200_I16) return tiles_trace<SegI16I16>(dev, (uint32_t)count, npipes, trace, cover, stats);
    if (intype == DOPPLER_B200_I16) return tiles_trace<SegI16F32>(dev, (uint32_t)count, npipes, trace, cover, stats);
    if (outtype == DOPPLER_B200_I16) return tiles_trace<SegF32I16>(dev, (uint32_t)count, npipes, trace, cover, stats);
    return tiles_trace<SegF32F32>(dev, (uint32_t)count, npipes, trace, cover, stats);
}

// ---- device self-test probes ------------------------------------------------------------------

int doppler_b200_phasor_probe(doppler_b200_ctx* ctx, float r, uint32_t n0, size_t count, float* cos_out, float* sin_out)
{
    if (!ctx || !cos_out || !sin_out) return DOPPLER_B200_EINVAL;
    if (count == 0) return DOPPLER_B200_OK;
    if (count > (1u << 30)) return fail(ctx, DOPPLER_B200_EINVAL, "probe too large");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    float *dc = nullptr, *ds = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&dc, count * 4));
    CUDA_TRY(ctx, cudaMalloc(&ds, count * 4));
    dmix::phasor_probe_kernel<<<(uint32_t)((count + 255) / 256), 256, 0, ctx->stream>>>(r, n0, (uint32_t)count, dc, ds);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(cos_out, dc, count * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sin_out, ds, count * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dc);
    cudaFree(ds);
    if (e != cudaSuccess) return fail(ctx, DOPPLER_B200_ECUDA, "phasor probe: %s", cudaGetErrorString(e));
    return DOPPLER_B200_OK;
}

int doppler_b200_sincosf_probe(doppler_b200_ctx* ctx, uint32_t first_bits, uint32_t stride, size_t count, float* sin_out,
                               float* cos_out)
{
    if (!ctx || !cos_out || !sin_out) return DOPPLER_B200_EINVAL;
    if (count == 0) return DOPPLER_B200_OK;
    if (count > (1u << 30)) return fail(ctx, DOPPLER_B200_EINVAL, "probe too large");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    float *dc = nullptr, *ds = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&dc, count * 4));
    CUDA_TRY(ctx, cudaMalloc(&ds, count * 4));
    dmix::sincosf_probe_kernel<<<(uint32_t)((count + 255) / 256), 256, 0, ctx->stream>>>(first_bits, stride, (uint32_t)count, ds, dc);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(cos_out, dc, count * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sin_out, ds, count * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dc);
    cudaFree(ds);
    if (e != cudaSuccess) return fail(ctx, DOPPLER_B200_ECUDA, "sincosf probe: %s", cudaGetErrorString(e));
    return DOPPLER_B200_OK;
}

}  // extern "C"
