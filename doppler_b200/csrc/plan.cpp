// plan.cpp -- see plan.h.  Host-only; built with -fno-fast-math -ffp-contract=off so that the
// f32 product r * f32(n) is the single IEEE multiply the reference performs (dsp.rs:125).
#include "plan.h"

#include <math.h>
#include <string.h>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

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

// Scalar scan (any ISA; also the tail / wrap-around path of the vector scans).
static uint64_t first_hit_scalar(float r, uint32_t n0, uint64_t limit)
{
    for (uint64_t i = 0; i < limit; i++)
        if (hit(r, n0 + (uint32_t)i)) return i;
    return limit;
}

#if defined(__x86_64__) && defined(__GNUC__)
// The same test on 8 / 16 consecutive n per step.  The planner runs this scan once per change of
// shift (track mode: once per second of stream) over up to one period; the auto-vectoriser does
// not vectorise hit() (u32 -> f32 conversion), which left the host planning of a 600-run schedule
// 10x slower than the kernels that mix it.  n < 2^31 here, so the signed conversion is exact.
__attribute__((target("avx2"))) static uint64_t first_hit_avx2(float r, uint32_t n0, uint64_t limit)
{
    const __m256 vr = _mm256_set1_ps(r), two23 = _mm256_set1_ps(8388608.0f), fmax = _mm256_set1_ps(3.40282347e+38f);
    const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
    __m256i vn = _mm256_add_epi32(_mm256_set1_epi32((int)n0), _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7));
    const __m256i step = _mm256_set1_epi32(8);
    uint64_t i = 0;
    for (; i + 8 <= limit; i += 8, vn = _mm256_add_epi32(vn, step)) {
        const __m256 x = _mm256_mul_ps(vr, _mm256_cvtepi32_ps(vn));
        const __m256 ax = _mm256_and_ps(x, absmask);
        const __m256 small = _mm256_cmp_ps(ax, two23, _CMP_LT_OQ);
        const __m256 back = _mm256_cvtepi32_ps(_mm256_cvttps_epi32(_mm256_and_ps(x, small)));
        const __m256 small_int = _mm256_and_ps(small, _mm256_cmp_ps(back, x, _CMP_EQ_OQ));
        const __m256 big = _mm256_and_ps(_mm256_cmp_ps(ax, two23, _CMP_GE_OQ), _mm256_cmp_ps(ax, fmax, _CMP_LE_OQ));
        const int m = _mm256_movemask_ps(_mm256_or_ps(small_int, big));
        if (m) return i + (uint64_t)__builtin_ctz((unsigned)m);
    }
    const uint64_t d = first_hit_scalar(r, n0 + (uint32_t)i, limit - i);
    return i + d;
}

__attribute__((target("avx512f"))) static uint64_t first_hit_avx512(float r, uint32_t n0, uint64_t limit)
{
    const __m512 vr = _mm512_set1_ps(r), two23 = _mm512_set1_ps(8388608.0f), fmax = _mm512_set1_ps(3.40282347e+38f);
    __m512i vn = _mm512_add_epi32(_mm512_set1_epi32((int)n0), _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15));
    const __m512i step = _mm512_set1_epi32(16);
    uint64_t i = 0;
    for (; i + 16 <= limit; i += 16, vn = _mm512_add_epi32(vn, step)) {
        const __m512 x = _mm512_mul_ps(vr, _mm512_cvtepi32_ps(vn));
        const __m512 ax = _mm512_abs_ps(x);
        const __mmask16 small = _mm512_cmp_ps_mask(ax, two23, _CMP_LT_OQ);
        const __m512 xs = _mm512_maskz_mov_ps(small, x);
        const __m512 back = _mm512_cvtepi32_ps(_mm512_cvttps_epi32(xs));
        const __mmask16 small_int = _mm512_mask_cmp_ps_mask(small, back, x, _CMP_EQ_OQ);
        const __mmask16 big = _mm512_mask_cmp_ps_mask(_mm512_cmp_ps_mask(ax, two23, _CMP_GE_OQ), ax, fmax, _CMP_LE_OQ);
        const unsigned m = (unsigned)(small_int | big);
        if (m) return i + (uint64_t)__builtin_ctz(m);
    }
    const uint64_t d = first_hit_scalar(r, n0 + (uint32_t)i, limit - i);
    return i + d;
}
#endif

// n in [2^31, 2^32): f32(n) takes one value per 256 consecutive n (ulp = 2^8, ties to the even mantissa), and the
// reset test depends on n only through f32(n): one test per plateau instead of one per n.  Without this a ratio
// that does not reset before the u32 wrap (tiny |r|) costs 2^31 scalar tests -- seconds of host time inside plan().
static uint64_t first_hit_plateaus(float r, uint32_t n0, uint64_t limit)
{
    uint64_t done = 0;
    while (done < limit) {
        const uint32_t n = n0 + (uint32_t)done;   // >= 2^31 by contract; n0 + limit <= 2^32
        if (hit(r, n)) return done;
        const uint64_t v = (uint64_t)(float)n;                          // the plateau's value: a multiple of 256, up to 2^32
        uint64_t last = v + (((v >> 8) & 1) ? 127 : 128);               // largest n rounding to v (the tie goes to the even mantissa)
        if (last > 0xffffffffull) last = 0xffffffffull;
        done += last + 1 - n;
    }
    return limit;
}

uint64_t first_hit(float r, uint32_t n0, uint64_t limit)
{
    // a non-finite ratio never hits (Inf * n is Inf or NaN, dsp.rs:125 fract() of those is NaN): no scan
    if (!(r - r == 0.0f)) return limit;
    uint64_t done = 0;
#if defined(__x86_64__) && defined(__GNUC__)
    static const int isa = __builtin_cpu_supports("avx512f") ? 2 : __builtin_cpu_supports("avx2") ? 1 : 0;
    // vector scans need n < 2^31 throughout (signed conversion); the stretch beyond is scalar
    while (isa && done < limit) {
        const uint32_t nb = n0 + (uint32_t)done;
        if (nb >= 0x80000000u) break;
        const uint64_t room = 0x80000000ull - nb, span = limit - done < room ? limit - done : room;
        const uint64_t d = isa == 2 ? first_hit_avx512(r, nb, span) : first_hit_avx2(r, nb, span);
        if (d < span) return done + d;
        done += span;
        if (span == room) break;   // continue scalar through the sign boundary and beyond
    }
#endif
    // scalar: no vector ISA, or n >= 2^31 (including the u32 wrap-around)
    while (done < limit) {
        const uint32_t nb = n0 + (uint32_t)done;
        uint64_t span = limit - done;
#if defined(__x86_64__) && defined(__GNUC__)
        if (isa && nb < 0x80000000u) break;   // wrapped back below 2^31: hand back to the vector scan
#endif
        const uint64_t room = 0x100000000ull - nb;   // up to the u32 wrap
        if (span > room) span = room;
        const uint64_t d = nb >= 0x80000000u ? first_hit_plateaus(r, nb, span) : first_hit_scalar(r, nb, span);
        if (d < span) return done + d;
        done += span;
    }
    if (done < limit) return done + first_hit(r, n0 + (uint32_t)done, limit - done);
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
    // Arbitrary start state (a shift change in track mode lands anywhere): scan, and remember the
    // answer -- a caller that replays or re-plans the same schedule pays for each scan once.
    const uint64_t key = ((uint64_t)fbits(r) << 32) | n;
    auto it = hits_.find(key);
    if (it != hits_.end()) {
        if (it->second.found) return it->second.dist < count ? it->second.dist : count;
        if (it->second.dist >= count) return count;   // dist = samples known to be hit-free
    }
    const uint64_t done = it != hits_.end() ? it->second.dist : 0;
    const uint64_t d = done + first_hit(r, n + (uint32_t)done, count - done);
    hits_[key] = HitInfo{d < count ? d : count, d < count};
    return d < count ? d : count;
}

void Planner::plan(const std::vector<Run>& runs, uint64_t k0, uint32_t* samplenum, std::vector<Piece>* out)
{
    // Realtime track mode forms a new ratio for every 8192-byte block, for hours: keep the per-ratio caches
    // bounded (trimmed here, where no reference into them is live).
    if (cache_.size() > (1u << 16)) cache_.clear();
    if (hits_.size() > (1u << 16)) hits_.clear();
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
    // Replay schedules repeat one shift for a whole second of blocks (millions of blocks at 200 Msps):
    // find the end of each stretch of identical shift bits by galloping with memcmp (the array against
    // itself one element on), then form the ratio once per stretch.
    const size_t need = (size_t)((total_samples + block_samples - 1) / block_samples);
    const size_t nb = need < nblocks ? need : nblocks;
    constexpr size_t kGallop = 1024;
    for (size_t b = 0; b < nb && left > 0;) {
        size_t e = b + 1;
        while (e + kGallop <= nb && memcmp(shift_hz + e - 1, shift_hz + e, kGallop * sizeof(float)) == 0) e += kGallop;
        while (e < nb && fbits(shift_hz[e]) == fbits(shift_hz[b])) e++;
        const uint64_t span = (uint64_t)(e - b) * block_samples;
        const uint64_t c = left < span ? left : span;
        const float r = ratio(shift_hz[b], samplerate);
        if (!runs.empty() && fbits(runs.back().r) == fbits(r))
            runs.back().count += c;
        else
            runs.push_back(Run{c, r});
        left -= c;
        b = e;
    }
    return runs;
}

}  // namespace dplan
