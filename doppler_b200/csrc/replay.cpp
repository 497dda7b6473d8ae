// replay.cpp -- host side of track mode's shift schedule.
//
// The reference's replay driver (/root/reference/src/main.rs:155-184) asks the orbit propagator
// for a Doppler value before every 8192-byte block, at `start_time + dt`, where dt is a whole
// number of seconds computed one iteration earlier from the samples counted before the previous
// block, in f32 (main.rs:166).  This file reproduces that clock exactly, with the propagator
// abstracted as a per-second table, so that the whole recording can be planned up front and
// mixed in large launches (doppler_b200_mix_blocks) with byte-identical results.
// Built with -ffp-contract=off -fno-fast-math (f32 arithmetic as rustc emits it).
#include <math.h>
#include <stdint.h>

#include "../../include/doppler_b200.h"

#ifdef __FAST_MATH__
#error "replay.cpp must not be built with -ffast-math"
#endif

extern "C" {

// main.rs:163  (range_rate_km_sec * 1000 / c) * frequency as f64 * (-1.0), evaluated in f64 in
// exactly this association order.
double doppler_b200_doppler_hz(double range_rate_km_sec, uint32_t frequency)
{
    const double speed_of_light_m_s = 299792458.0;   // main.rs:48
    return (range_rate_km_sec * 1000.0 / speed_of_light_m_s) * (double)frequency * (-1.0);
}

// main.rs:177  doppler_hz as f32 + offset as f32
float doppler_b200_track_shift(double doppler_hz, int32_t offset) { return (float)doppler_hz + (float)offset; }

// main.rs:166  whole seconds of stream time after `sample_count` samples: (n as f32 / fs as f32) as i64.
// Rust's float -> int `as` saturates and maps NaN to 0 (fs == 0 gives +inf or NaN).
int64_t doppler_b200_replay_seconds(uint64_t sample_count, uint32_t samplerate)
{
    const float t = (float)sample_count / (float)samplerate;
    if (t != t) return 0;
    if (t >= 9223372036854775807.0f) return INT64_MAX;
    if (t <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)t;
}

size_t doppler_b200_replay_schedule(const double* doppler_hz_by_second, size_t nsec, int32_t offset, uint32_t samplerate,
                                    int intype, size_t in_len, float* shifts_out, size_t cap)
{
    if (!doppler_hz_by_second || nsec == 0 || (intype != DOPPLER_B200_I16 && intype != DOPPLER_B200_F32)) return 0;
    const size_t bps = intype == DOPPLER_B200_I16 ? 4 : 8;
    const size_t block_samples = DOPPLER_B200_BUFFER_SIZE / bps;
    // the pump stops on the first short read (main.rs:98,178): one more block than the full ones
    const size_t nblocks = in_len / DOPPLER_B200_BUFFER_SIZE + 1;
    uint64_t sample_count = 0;   // main.rs:157
    int64_t dt = 0;              // main.rs:158
    for (size_t b = 0; b < nblocks; b++) {
        const size_t idx = dt < 0 ? 0 : ((uint64_t)dt < nsec ? (size_t)dt : nsec - 1);
        const double doppler_hz = doppler_hz_by_second[idx];             // main.rs:162-163 at start_time + dt
        dt = doppler_b200_replay_seconds(sample_count, samplerate);      // main.rs:166 (used by the NEXT block)
        if (b < cap && shifts_out) shifts_out[b] = doppler_b200_track_shift(doppler_hz, offset);
        sample_count += block_samples;                                    // main.rs:182 (full blocks; the short one ends the loop)
    }
    return nblocks;
}

}  // extern "C"
