// plan.h -- host-side analytic planner for the reference's `samplenum` state machine.
//
// The reference mixer is serially dependent through one u32 (src/dsp.rs:125-130, state at
// src/main.rs:60): after every sample, samplenum <- 1 if fract(r * f32(samplenum)) == 0 else
// samplenum + 1, with r = shift_hz / f32(samplerate) evaluated in f32.  The phase index is
// therefore NOT the absolute sample index.  For a constant r the sequence is piecewise simple:
// it counts up from its start value until the first "hit" (reset), then repeats 1..P forever,
// P being the smallest n >= 1 with a hit.  The planner turns a list of constant-shift runs
// into PIECES in which samplenum is a closed-form function of the sample index, so that any
// GPU thread (or any GPU of a time-sliced job) can compute its own samplenum independently:
//
//   linear   piece (period == 0): samplenum(k) = base + (k - k_begin)            (u32 wrap)
//   periodic piece (period  > 0): samplenum(k) = ((base + (k - k_begin)) mod period) + 1
//
// Pure host code (no CUDA): unit-tested against the oracle's sequential recurrence.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <unordered_map>
#include <vector>

namespace dplan {

// Largest period represented as a periodic piece; longer periods become linear pieces (one
// per period), so base + offset always fits 32 bits inside one kernel launch (<= 2^30 samples).
constexpr uint64_t kPeriodicMax = 1ull << 30;

struct Run {        // `count` consecutive samples mixed with one shift value
    uint64_t count;
    float r;        // shift_hz / (float)samplerate, computed in f32 exactly as dsp.rs:121
};

struct Piece {      // stream-absolute sample indices
    uint64_t k_begin;
    uint64_t k_end;
    uint32_t base;
    uint32_t period;   // 0 -> linear
    float r;
};

// dsp.rs:121/125: r = shift_hz / samplerate as f32
float ratio(float shift_hz, uint32_t samplerate);

// dsp.rs:125: fract(r * n as f32) == 0.0
bool reset_test(float r, uint32_t n);

// Smallest d in [0, limit) such that reset_test(r, n0 + d) (u32 wrap-around); `limit` if none.
uint64_t first_hit(float r, uint32_t n0, uint64_t limit);

class Planner {
public:
    // Appends the pieces covering `runs` (starting at stream index k0 with state *samplenum)
    // to `out` and advances *samplenum to the state after the last sample.
    void plan(const std::vector<Run>& runs, uint64_t k0, uint32_t* samplenum, std::vector<Piece>* out);

    // State only (no piece list): what the reference's samplenum is after the runs.
    uint32_t advance(const std::vector<Run>& runs, uint32_t samplenum);

    // Smallest P >= 1 with reset_test(r, P), if P <= limit; 0 if there is none up to `limit`.
    // (P can be 2^32: the u32 wraps to 0, and r*0 is a hit for every finite r.)
    uint64_t period(float r, uint64_t limit);

private:
    struct PeriodInfo {
        uint64_t searched = 0;   // n in [1, searched] are known not to hit (when period == 0)
        uint64_t period = 0;
    };
    struct HitInfo {
        uint64_t dist;   // found: samples until the first hit; otherwise: samples known to be hit-free
        bool found;
    };
    std::unordered_map<uint32_t, PeriodInfo> cache_;   // key: bit pattern of r
    std::unordered_map<uint64_t, HitInfo> hits_;       // key: (bits of r, start samplenum) -> first_hit result
    const PeriodInfo& learn(float r, uint32_t n, uint64_t count);
    uint64_t hit_distance(float r, uint32_t n, uint64_t count);
};

// Groups per-block shifts (src/main.rs:177: one shift per BUFFER_SIZE-byte block) into runs.
// total_samples may end inside the last block.
std::vector<Run> runs_from_blocks(const float* shift_hz, size_t nblocks, uint64_t block_samples,
                                  uint32_t samplerate, uint64_t total_samples);

}  // namespace dplan
