"""The analytic samplenum planner (doppler_b200/csrc/plan.cpp) against the oracle's sequential
recurrence (restating /root/reference/src/dsp.rs:125-130).  Host logic only -- no GPU."""
import numpy as np
import pytest

from doppler_b200 import dsp

# (shift_hz, samplerate): regular ratios, irregular ratios with large periods, integers, zero,
# |r| > 1, tiny r (no reset), negative
CASES = [
    (-15000.0, 256000), (100000.0, 10_000_000), (815000.0, 2_400_000), (5000.0, 1_024_000),
    (-9876.54, 1_024_000), (7321.7, 1_024_000), (7321.0, 1_024_000), (0.0, 48000), (48000.0, 48000),
    (-96000.0, 48000), (123456.0, 1000), (1.0, 2_000_000_000), (4_000_000.5, 200_000_000),
    (-3_912_345.25, 200_000_000), (-117_187_500.0, 2_000_000_000), (0.001, 48000), (3.0e38, 1),
    (float("inf"), 48000), (float("nan"), 48000), (1000.0, 0),
]


@pytest.mark.parametrize("shift,fs", CASES)
def test_trace_single_run(oracle, shift, fs):
    count = 300_000
    want, sn_want = oracle.samplenum_trace(0, shift, fs, count)
    got, sn, npieces = dsp.plan_trace(0, [shift], count, fs, count)
    assert np.array_equal(got, want)
    assert sn == sn_want
    assert npieces <= 4 + count // 60_000  # closed form, not one piece per period


@pytest.mark.parametrize("start", [0, 1, 2, 255, 256, 257, 100_000, 2**24 + 3, 2**32 - 5, 2**32 - 1])
def test_trace_arbitrary_start_state(oracle, start):
    """samplenum handed in mid-stream (library callers carry it, main.rs:60), including the u32 wrap."""
    for shift, fs in [(-15000.0, 256000), (7321.7, 1_024_000), (1.0, 2_000_000_000)]:
        want, sn_want = oracle.samplenum_trace(start, shift, fs, 5000)
        got, sn, _ = dsp.plan_trace(start, [shift], 5000, fs, 5000)
        assert np.array_equal(got, want), (start, shift, fs)
        assert sn == sn_want


def test_trace_block_schedule_carries_state_across_shift_changes(oracle):
    """Track mode: a new shift per block, samplenum NOT reset (main.rs:60,177)."""
    rng = np.random.default_rng(5)
    fs = 1_024_000
    block = 2048
    shifts = np.concatenate([
        np.repeat(np.float32(-9876.54), 40), np.repeat(np.float32(-9871.02), 37), rng.uniform(-12000, 12000, 30).astype(np.float32),
        np.repeat(np.float32(5000.0), 25), np.repeat(np.float32(0.0), 3), np.repeat(np.float32(7321.7), 64),
    ])
    count = shifts.size * block - 777  # last block short
    want = np.empty(count, dtype=np.uint32)
    sn = 0
    k = 0
    for s in shifts:
        n = min(block, count - k)
        tr, sn = oracle.samplenum_trace(sn, float(s), fs, n)
        want[k:k + n] = tr
        k += n
    got, sn_got, npieces = dsp.plan_trace(0, shifts, block, fs, count)
    assert np.array_equal(got, want)
    assert sn_got == sn
    assert dsp.samplenum_advance_blocks(0, shifts, block, fs, count) == sn


@pytest.mark.parametrize("shift,fs", [(-15000.0, 256000), (100000.0, 10_000_000), (7321.7, 1_024_000), (-9876.54, 1_024_000),
                                      (1.0, 2_000_000_000), (0.0, 1000)])
def test_advance_matches_oracle(oracle, shift, fs):
    for start in (0, 1, 77):
        for count in (0, 1, 2, 255, 256, 257, 99_999, 1_000_003):
            assert dsp.samplenum_advance(start, shift, fs, count) == oracle.samplenum_advance(start, shift, fs, count)


def test_advance_is_analytic_for_long_streams(oracle):
    """Time-slice seeds for a 200 Msps x 60 s stream cut 8 ways (cfg4) must not cost O(count)."""
    import time
    fs = 200_000_000
    t0 = time.time()
    seeds = [dsp.samplenum_advance(0, 4_000_000.5, fs, i * 1_500_000_000) for i in range(8)]
    assert time.time() - t0 < 5.0
    # chain property: advancing slice by slice equals advancing in one go
    sn = 0
    for i in range(1, 8):
        sn = dsp.samplenum_advance(sn, 4_000_000.5, fs, 1_500_000_000)
        assert sn == seeds[i]
    # spot-check one seed against the sequential oracle on a shorter prefix
    assert dsp.samplenum_advance(0, 4_000_000.5, fs, 30_000_000) == oracle.samplenum_advance(0, 4_000_000.5, fs, 30_000_000)


def test_slices_compose(oracle):
    """(e) multi-GPU: any cut of the stream into contiguous slices seeded analytically reproduces the whole trace."""
    fs, shift, count = 1_024_000, -9876.54, 500_000
    whole, _ = oracle.samplenum_trace(0, shift, fs, count)
    cuts = [0, 8192, 131072, 300_000, count]
    for a, b in zip(cuts[:-1], cuts[1:]):
        seed = dsp.samplenum_advance(0, shift, fs, a)
        part, _, _ = dsp.plan_trace(seed, [shift], b - a, fs, b - a)
        assert np.array_equal(part, whole[a:b])


# ---- track mode's Doppler schedule (replay.cpp) against the oracle's replay driver -------------

@pytest.mark.parametrize("intype", [0, 1])
@pytest.mark.parametrize("fs", [8192, 256000, 1_024_000, 3])
def test_replay_schedule_matches_reference_clock(oracle, intype, fs):
    """main.rs:155-184: per-block f32 shift with the one-block lag and f32 second arithmetic."""
    t = np.arange(40, dtype=np.float64)
    rr = 7.5 * 7.5 * (t - 17.0) / np.sqrt(700.0 ** 2 + (7.5 * (t - 17.0)) ** 2)
    table = np.array([oracle.doppler_hz(x, 437_505_000) for x in rr])
    assert [dsp.doppler_hz(x, 437_505_000) for x in rr] == list(table)
    bps = 4 if intype == 0 else 8
    for nbytes in (0, bps, 8192, 8192 * 7 + bps * 3, 8192 * 40, 8192 * 133 + bps):
        x = np.zeros(nbytes, dtype=np.uint8)
        _, _, used, panicked = oracle.track_replay_stream(x, intype, intype, table, -2500, fs)
        assert not panicked
        got = dsp.replay_schedule(table, -2500, fs, intype, nbytes)
        assert np.array_equal(used.view(np.uint32), got.view(np.uint32)), (nbytes, fs)


# ---- work decomposition of one kernel launch (GRID / COLUMN segments, mixer_kernels.cuh) -------------
TILE_CASES = [
    # (shift, fs, start samplenum, samples): periods above the shared-memory table size become COLUMN segments
    (-9876.54, 1_024_000, 0, 5_000_001),          # P = 111 145 (odd): every row has a different alignment shift
    (7321.7, 1_024_000, 0, 4_224_003),            # P = 55 244
    (-3_912_345.25, 200_000_000, 12_345, 4_500_000),  # P = 26 787, starts mid-period
    (7321.7, 1_024_000, 0, 1_024_003),            # below the COLUMN launch threshold: GRID tiles, direct evaluation
    (4_000_000.5, 200_000_000, 0, 400_000),       # P = 4.9 M > samples: linear piece only, GRID
    (-15000.0, 256000, 7, 100_003),               # P = 256: table piece, GRID only
    (5000.0, 1_024_000, 0, 300_000),              # P = 1024
    (0.0, 48000, 0, 10_001), (1.0, 2_000_000_000, 2**32 - 1000, 50_000),
]


@pytest.mark.parametrize("intype,outtype", [(dsp.I16, dsp.I16), (dsp.I16, dsp.F32), (dsp.F32, dsp.I16), (dsp.F32, dsp.F32)])
@pytest.mark.parametrize("shift,fs,start,count", TILE_CASES)
def test_launch_tiles_cover_every_sample_once_with_the_reference_samplenum(oracle, intype, outtype, shift, fs, start, count):
    """The segment builder and the kernel's own tile iterator, walked on the host: every sample below the
    sub-granule tail is covered by exactly one tile, and the samplenum the tile arithmetic assigns (COLUMN:
    window entry of the parked phasors; GRID: closed-form piece) is the reference recurrence's."""
    want, _ = oracle.samplenum_trace(start, shift, fs, count)
    for npipes in (148 * 20, 7):
        trace, cover, tail, stats = dsp.plan_tiles_trace(intype, outtype, start, [shift], count, fs, count, npipes)
        assert count - tail < 4
        assert np.all(cover[:tail] == 1) and np.all(cover[tail:] == 0), stats
        assert np.array_equal(trace[:tail], want[:tail]), stats
        if count >= 4 << 20 and abs(shift) in (9876.54, 7321.7, 3_912_345.25):
            assert stats["column_segments"] == 1, stats   # long period, several periods, large launch: the COLUMN path


def test_launch_tiles_track_schedule(oracle):
    """A per-block schedule (track mode): several periodic pieces with long periods in one launch."""
    fs, block = 1_024_000, 2048
    shifts = np.concatenate([np.repeat(np.float32(-9876.54), 700), np.repeat(np.float32(7321.7), 400),
                             np.repeat(np.float32(5000.0), 100), np.repeat(np.float32(-3211.11), 900)])
    count = shifts.size * block - 333
    want = np.empty(count, dtype=np.uint32)
    sn, k = 0, 0
    for s in (-9876.54, 7321.7, 5000.0, -3211.11):
        n = min(int((shifts == np.float32(s)).sum()) * block, count - k)
        want[k:k + n], sn = oracle.samplenum_trace(sn, float(np.float32(s)), fs, n)
        k += n
    for it, ot in [(dsp.I16, dsp.I16), (dsp.F32, dsp.F32)]:
        trace, cover, tail, stats = dsp.plan_tiles_trace(it, ot, 0, shifts, block, fs, count)
        assert stats["column_segments"] >= 2, stats
        assert np.all(cover[:tail] == 1) and np.array_equal(trace[:tail], want[:tail])


def test_launch_tiles_random_schedules(oracle):
    """Randomised schedules (run lengths from one block to a few thousand, shifts drawn from values with short,
    long and no reset periods, arbitrary start state and ragged ends): coverage and samplenum of the launch's
    tiles against the planner's own trace (itself checked against the oracle above)."""
    rng = np.random.default_rng(20151122)
    pool = np.array([-9876.54, 7321.7, 5000.0, -3211.11, -15000.0, 0.0, 12_345.678, 1.0, 815000.0, -1234.5, 48000.0], dtype=np.float32)
    for trial in range(24):
        it, ot = [(dsp.I16, dsp.I16), (dsp.I16, dsp.F32), (dsp.F32, dsp.I16), (dsp.F32, dsp.F32)][trial % 4]
        fs = int(rng.choice([96_000, 1_024_000, 2_400_000]))
        block = 2048 if it == dsp.I16 else 1024
        nruns = int(rng.integers(1, 7))
        shifts = np.concatenate([np.repeat(rng.choice(pool), int(rng.integers(1, 1500))) for _ in range(nruns)])
        count = int(shifts.size * block - rng.integers(0, block))
        start = int(rng.choice([0, 1, 77_777, 2**24 + 5, 2**32 - 3]))
        npipes = int(rng.choice([1, 13, 148 * 12]))
        want, _, _ = dsp.plan_trace(start, shifts, block, fs, count)
        trace, cover, tail, stats = dsp.plan_tiles_trace(it, ot, start, shifts, block, fs, count, npipes)
        assert count - tail < 4, (trial, stats)
        assert np.all(cover[:tail] == 1) and np.all(cover[tail:] == 0), (trial, stats)
        assert np.array_equal(trace[:tail], want[:tail]), (trial, stats)


def test_trace_property_random_ratios(oracle):
    """Property: for ANY (shift, samplerate, start state) the planner's closed-form pieces expand to the
    reference recurrence (dsp.rs:125-130) sample for sample, and the carried state matches."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(shift=st.one_of(st.integers(-2_000_000, 2_000_000).map(float), st.floats(-1e6, 1e6, allow_nan=False, width=32)),
           fs=st.sampled_from([8000, 48000, 96000, 256000, 1_024_000, 2_400_000, 10_000_000, 200_000_000]),
           start=st.one_of(st.integers(0, 300_000), st.integers(2**24 - 5, 2**24 + 5), st.integers(2**32 - 20_000, 2**32 - 1)),
           count=st.integers(1, 40_000))
    def prop(shift, fs, start, count):
        want, sn_want = oracle.samplenum_trace(start, shift, fs, count)
        got, sn, _ = dsp.plan_trace(start, [shift], count, fs, count)
        assert np.array_equal(got, want)
        assert sn == sn_want
        assert dsp.samplenum_advance(start, shift, fs, count) == sn_want

    prop()


def test_upper_half_plateau_scan_matches_the_oracle(oracle):
    """n >= 2^31: the planner tests one n per f32 plateau (256 wide, ties to even) -- same answers as the recurrence."""
    fs = 1_000_000
    r_hits_late = float(np.float32(2.0 ** -10 * (1 + 2.0 ** -20)) * np.float32(fs))   # r*f32(n) ~ 2^21: integer on ~1 plateau in 4
    for shift in (r_hits_late, 1.0e-6, -3.3e-4):
        for start in (2**31, 2**31 + 12345, 2**31 + 127, 2**31 + 128, 2**31 + 129, 2**32 - 70_000, 3 * 2**30 + 383):
            want, sn_want = oracle.samplenum_trace(start, shift, fs, 100_000)
            got, sn, _ = dsp.plan_trace(start, [shift], 100_000, fs, 100_000)
            assert np.array_equal(got, want), (shift, start)
            assert sn == sn_want


def test_no_reset_before_the_wrap_is_planned_quickly():
    """A ratio that cannot reset before samplenum wraps (|r| * 2^32 < 1): 2^31 scalar tests before, one per plateau now."""
    import time
    t0 = time.perf_counter()
    sn = dsp.samplenum_advance(2**31, 1.0e-7, 4_000_000_000, 2**31 + 10)   # r = 2.5e-17
    dt = time.perf_counter() - t0
    assert sn == 10          # ... 2^32 wraps to 0, r*0 == 0 resets to 1, nine more samples
    assert dt < 2.0, dt
    t0 = time.perf_counter()
    for shift in (float("inf"), float("nan")):
        assert dsp.samplenum_advance(5, shift, 48000, 2**33) == (5 + 2**33) % 2**32   # never resets, never scans
    assert time.perf_counter() - t0 < 0.5


def test_libm_guard(oracle):
    """The numerics guard: the host twin of the device sincosf against this host's libm (compatible here), and
    against a deliberately different sincosf (detected)."""
    import ctypes
    import math
    from doppler_b200 import _lib
    lib = _lib.load()
    assert dsp.libm_compatible()
    assert lib.doppler_b200_libm_mismatches(None) == 0
    CB = ctypes.CFUNCTYPE(None, ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float))

    def other(y, s, c):   # double-precision sin/cos rounded to float: close, but not glibc's sincosf
        s[0] = math.sin(y) if math.isfinite(y) else float("nan")
        c[0] = math.cos(y) if math.isfinite(y) else float("nan")

    cb = CB(other)
    assert lib.doppler_b200_libm_mismatches(ctypes.cast(cb, ctypes.c_void_p)) > 0


def test_steady_state_rule_of_the_per_block_path(oracle):
    """tiny_host_call (doppler_b200.cu) plans a block of an unchanged ratio without the planner when samplenum is inside the
    ratio's period P: ONE periodic piece with base = samplenum - 1, i.e. sample k uses ((samplenum - 1 + k) mod P) + 1, and
    the state afterwards is ((samplenum - 1 + n) mod P) + 1.  Pinned here against the oracle's sequential recurrence and the
    planner, for regular and irregular ratios, every kind of start inside the period and ragged block lengths."""
    rng = np.random.default_rng(2026)
    for shift, fs in [(5000.0, 1_024_000), (-15000.0, 256000), (100000.0, 10_000_000), (7321.7, 1_024_000), (-9876.54, 1_024_000)]:
        # the period: the state right after the first reset counts 1 .. P
        first, _ = oracle.samplenum_trace(1, shift, fs, 400_000)
        resets = np.flatnonzero(first[1:] == 1)
        assert resets.size, (shift, fs)
        P = int(first[resets[0]])                        # the value that hit
        starts = [1, 2, P - 1, P] + [int(x) for x in rng.integers(1, P + 1, 6)]
        for sn in starts:
            for n in (1, 3, 2047, 2048, 8192, int(rng.integers(1, 3 * P + 10))):
                want, sn_want = oracle.samplenum_trace(sn, shift, fs, n)
                k = np.arange(n, dtype=np.uint64)
                rule = ((np.uint64(sn - 1) + k) % np.uint64(P) + 1).astype(np.uint32)
                assert np.array_equal(rule, want), (shift, fs, sn, n)
                assert ((sn - 1 + n) % P) + 1 == sn_want
                got, sn_got, npieces = dsp.plan_trace(sn, [shift], n, fs, n)
                assert np.array_equal(got, want) and sn_got == sn_want
                # (the planner has to learn the period first: it may need a piece per period until a stream that started
                #  at a reset has shown it -- the per-block path takes its shortcut only from a plan that was one periodic piece)
