// plan.cpp -- see plan.h.  Host-only; built with -fno-fast-math -ffp-contract=off so that the
// f32 product r * f32(n) is the single IEEE multiply the reference performs (dsp.rs:125).
#include "plan.h"

#include <math.h>
#include <string.h>

#ifdef __FAST_MATH__
#error "plan.cpp must not be built with -ffast-math"
#endif

namespace dplan {

static inline uint32_t fbits(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

float ratio(float shift_hz, uint32_t samplerate) { return shift_hz / (float)samplerate; }

// fract(x) == 0  <=>  x finite and integer-valued.  Written without truncf so that the search
// loop vectorises on plain SSE2: for |x| < 2^23 compare against the int round trip, every
// finite |x| >= 2^23 is an integer, Inf/NaN never hit (Inf - Inf = NaN in f32::fract).
static inline bool hit(float r, uint32_t n)
{
    const float x = r * (float)n;
    const float ax = fabsf(x);
    const bool small = ax < 8388608.0f;
    const float xc = small ? x : 0.0f;   // keeps the int conversion in range (blend, still vectorises)
    const bool small_int = small & ((float)(int32_t)xc == xc);
    const bool big = (ax >= 8388608.0f) & (ax <= 3.40282347e+38f);
    return small_int | big;
}

bool reset_test(float r, uint32_t n) { return hit(r, n); }

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
uint64_t first_hit(float r, uint32_t n0, uint64_t limit)
{
    constexpr uint64_t kChunk = 4096;
    for (uint64_t done = 0; done < limit; done += kChunk) {
        const uint64_t m = limit - done < kChunk ? limit - done : kChunk;
        const uint32_t nb = n0 + (uint32_t)done;
        unsigned any = 0;
        for (uint64_t i = 0; i < m; i++) any |= (unsigned)hit(r, nb + (uint32_t)i);
        if (any) {
            for (uint64_t i = 0; i < m; i++)
                if (hit(r, nb + (uint32_t)i)) return done + i;
        }
    }
    return limit;
}

uint64_t Planner::period(float r, uint64_t limit)
{
    PeriodInfo& pi = cache_[fbits(r)];
    if (pi.period) return pi.period <= limit ? pi.period : 0;
    if (limit > (1ull << 32)) limit = 1ull << 32;
    if (pi.searched >= limit) return 0;
    // continue the search at n = searched + 1 (n = 2^32 is the wrap to 0)
    const uint64_t span = limit - pi.searched;
    const uint64_t d = first_hit(r, (uint32_t)(pi.searched + 1), span);
    if (d < span) {
        pi.period = pi.searched + 1 + d;
        return pi.period;
    }
    pi.searched = limit;
    return 0;
}

// Makes the cache know about hits in [1, n + count - 1] whenever that range is contiguous with
// what has been searched already (it always is for a stream that started at 0 or at a reset).
const Planner::PeriodInfo& Planner::learn(float r, uint32_t n, uint64_t count)
{
    PeriodInfo& pi = cache_[fbits(r)];
    if (!pi.period && n >= 1 && (uint64_t)n <= pi.searched + 1 && (uint64_t)n + count - 1 > pi.searched)
        period(r, (uint64_t)n + count - 1);
    return pi;
}

// Samples until the next hit when starting in state n: d in [0, count) such that the sample
// d positions ahead uses a hitting samplenum, or `count` if no hit occurs within the run.
uint64_t Planner::hit_distance(float r, uint32_t n, uint64_t count)
{
    if (n >= 1) {
        const PeriodInfo& pi = learn(r, n, count);
        if (pi.period && n <= pi.period) {
            const uint64_t d = pi.period - n;
            return d < count ? d : count;
        }
        if (!pi.period && (uint64_t)n + count - 1 <= pi.searched) return count;
    }
    return first_hit(r, n, count);
}

void Planner::plan(const std::vector<Run>& runs, uint64_t k0, uint32_t* samplenum, std::vector<Piece>* out)
{
    uint32_t n = *samplenum;
    uint64_t k = k0;
    for (const Run& run : runs) {
        uint64_t count = run.count;
        const float r = run.r;
        while (count > 0) {
            if (n >= 1) {
                const PeriodInfo& pi = learn(r, n, count);
                if (pi.period && n <= pi.period && pi.period <= kPeriodicMax) {
                    const uint64_t P = pi.period;
                    if (out) out->push_back(Piece{k, k + count, n - 1, (uint32_t)P, r});
                    n = (uint32_t)(((uint64_t)(n - 1) + count) % P) + 1;
                    k += count;
                    count = 0;
                    break;
                }
            }
            const uint64_t d = hit_distance(r, n, count);
            if (d >= count) {
                if (out) out->push_back(Piece{k, k + count, n, 0u, r});
                n += (uint32_t)count;
                k += count;
                count = 0;
            } else {
                if (out) out->push_back(Piece{k, k + d + 1, n, 0u, r});
                k += d + 1;
                count -= d + 1;
                n = 1;
            }
        }
    }
    *samplenum = n;
}

uint32_t Planner::advance(const std::vector<Run>& runs, uint32_t samplenum)
{
    plan(runs, 0, &samplenum, nullptr);
    return samplenum;
}

std::vector<Run> runs_from_blocks(const float* shift_hz, size_t nblocks, uint64_t block_samples,
                                  uint32_t samplerate, uint64_t total_samples)
{
    std::vector<Run> runs;
    uint64_t left = total_samples;
    for (size_t b = 0; b < nblocks && left > 0; b++) {
        const uint64_t c = left < block_samples ? left : block_samples;
        const float r = ratio(shift_hz[b], samplerate);
        if (!runs.empty() && fbits(runs.back().r) == fbits(r))
            runs.back().count += c;
        else
            runs.push_back(Run{c, r});
        left -= c;
    }
    return runs;
}

}  // namespace dplan
