"""Parity of the sm_100a mixer (through the C ABI) with the CPU oracle: bit-exact bytes for i16
output, bit-exact floats for f32 output (NaN payloads aside).  Runs on the GPU box only.

The oracle is the restatement in oracle/ calling the reference's own compiled complex.c when
oracle/_ref is present (built in the authoring container, shipped with the snapshot)."""
import numpy as np
import pytest

import doppler_b200
from doppler_b200 import F32, I16
from tests.oracle_lib import BUFFER_SIZE, same_bits_f32

pytestmark = pytest.mark.gpu

BPS = {I16: 4, F32: 8}


def make_input(rng, n, typ, kind="uniform"):
    if typ == I16:
        if kind == "fullscale":
            v = rng.choice(np.array([-32768, -32767, 32767, 23170, -23170, 0, 1, -1], dtype=np.int16), 2 * n)
        else:
            v = rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype(np.int16)
        return v.view(np.uint8)
    v = rng.uniform(-0.7, 0.7, 2 * n).astype(np.float32)
    if kind == "fullscale":
        v = rng.choice(np.array([1.0, -1.0, 0.70710678, -0.70710678, 1.5, -2.0, 0.0, -0.0], dtype=np.float32), 2 * n)
    return v.view(np.uint8)


def check(oracle, got, want, outtype):
    assert got.size == want.size
    if outtype == I16:
        if not np.array_equal(got, want):
            g, w = got.view("<i2"), want.view("<i2")
            bad = np.flatnonzero(g != w)
            raise AssertionError(f"{bad.size} of {g.size} i16 values differ; first at {bad[0]}: got {g[bad[0]]} want {w[bad[0]]}")
    else:
        if not same_bits_f32(got, want):
            g, w = got.view(np.uint32), want.view(np.uint32)
            bad = np.flatnonzero(g != w)
            raise AssertionError(f"{bad.size} of {g.size} f32 words differ; first at {bad[0]}: got {g[bad[0]]:08x} want {w[bad[0]]:08x}")


TYPE_PAIRS = [(I16, I16), (I16, F32), (F32, I16), (F32, F32)]
# (shift, fs): table-in-shared-memory periods, an L2-sized table, direct evaluation, degenerate ratios
SHIFTS = [(-15000.0, 256000), (100000.0, 10_000_000), (815000.0, 2_400_000), (-9876.54, 1_024_000),
          (7321.7, 1_024_000), (0.0, 48000), (48000.0, 48000), (1.0, 2_000_000_000)]


@pytest.mark.parametrize("intype,outtype", TYPE_PAIRS)
@pytest.mark.parametrize("shift,fs", SHIFTS)
def test_mix_matches_oracle(oracle, mixer, intype, outtype, shift, fs):
    rng = np.random.default_rng(hash((intype, outtype, fs)) & 0xFFFF)
    for n in (1, 3, 2047, 4096, 4097, 70_001, 400_003):
        buf = make_input(rng, n, intype)
        got, sn = mixer.mix(buf, intype, outtype, shift, fs)
        want, sn_ref = oracle.mix(buf, intype, outtype, shift, fs)
        assert sn == sn_ref
        check(oracle, got, want, outtype)


@pytest.mark.parametrize("intype,outtype", TYPE_PAIRS)
def test_empty_and_misaligned_inputs(oracle, mixer, intype, outtype):
    got, sn = mixer.mix(np.zeros(0, dtype=np.uint8), intype, outtype, 1000.0, 48000, samplenum=17)
    assert got.size == 0 and sn == 17
    with pytest.raises(doppler_b200.DopplerError) as ei:  # the reference's assert!, dsp.rs:87,103
        mixer.mix(np.zeros(BPS[intype] * 5 + 2, dtype=np.uint8), intype, outtype, 1000.0, 48000)
    assert ei.value.code == doppler_b200.dsp.EALIGN


@pytest.mark.parametrize("intype,outtype", TYPE_PAIRS)
def test_fullscale_saturation(oracle, mixer, intype, outtype):
    """(1+1j)*e^{j theta} exceeds i16 range: Rust's saturating `as i16` (main.rs:77-78)."""
    rng = np.random.default_rng(99)
    buf = make_input(rng, 50_000, intype, "fullscale")
    got, _ = mixer.mix(buf, intype, outtype, 815000.0, 2_400_000)
    want, _ = oracle.mix(buf, intype, outtype, 815000.0, 2_400_000)
    check(oracle, got, want, outtype)


def test_f32_specials_pass_through(oracle, mixer):
    """NaN / Inf / denormal inputs (the f32 ingest is a bit copy, dsp.rs:101-115)."""
    pat = np.array([0x7FC00000, 0x3F800000, 0x7F800000, 0x00000000, 0xFF800000, 0x3F000000, 0x00000001, 0x80000001,
                    0x00800000, 0x7F7FFFFF, 0x3F800000, 0x7F7FFFFF], dtype=np.uint32)
    buf = np.tile(pat, 1000).view(np.uint8)
    for outtype in (I16, F32):
        got, _ = mixer.mix(buf, F32, outtype, 100000.0, 10_000_000)
        want, _ = oracle.mix(buf, F32, outtype, 100000.0, 10_000_000)
        check(oracle, got, want, outtype)


def test_samplenum_carried_across_calls(oracle, mixer):
    """Library callers pump block after block carrying samplenum (main.rs:60): chunked == whole."""
    rng = np.random.default_rng(4)
    n = 300_000
    buf = make_input(rng, n, I16)
    for shift, fs in [(-15000.0, 256000), (-9876.54, 1_024_000)]:
        want, sn_ref = oracle.mix(buf, I16, I16, shift, fs)
        sn = 0
        parts = []
        k = 0
        for m in (2048, 2048, 1, 99_999, 4096, n):  # last one clipped
            m = min(m, n - k)
            out, sn = mixer.mix(buf[4 * k:4 * (k + m)], I16, I16, shift, fs, samplenum=sn)
            parts.append(out)
            k += m
        assert k == n and sn == sn_ref
        check(oracle, np.concatenate(parts), want, I16)


def test_reference_named_functions(oracle, mixer):
    """dsp::convert_iqi16_to_complex / convert_iqf32_to_complex / shift_frequency one to one."""
    rng = np.random.default_rng(11)
    raw = rng.integers(-32768, 32768, 2 * 10_001, dtype=np.int32).astype("<i2")
    assert np.array_equal(mixer.convert_iqi16_to_complex(raw.tobytes()).view(np.uint32),
                          oracle.convert_iqi16_to_complex(raw.tobytes()).view(np.uint32))
    f = rng.uniform(-2, 2, 2 * 9_999).astype("<f4")
    f[:4] = [np.nan, np.inf, -0.0, 1e-42]
    assert np.array_equal(mixer.convert_iqf32_to_complex(f.tobytes()).view(np.uint32),
                          oracle.convert_iqf32_to_complex(f.tobytes()).view(np.uint32))
    with pytest.raises(doppler_b200.DopplerError):
        mixer.convert_iqi16_to_complex(b"\0" * 6)
    x = (rng.uniform(-1, 1, 125_000) + 1j * rng.uniform(-1, 1, 125_000)).astype(np.complex64)
    sn_g = sn_o = 0
    for _ in range(3):  # test_bench_shift_frequency's shape: 125 000 samples, 815 kHz @ 2.4 Msps (dsp.rs:136-157)
        got, sn_g = mixer.shift_frequency(x, sn_g, 815000.0, 2_400_000)
        want, sn_o = oracle.shift_frequency(x, sn_o, 815000.0, 2_400_000)
        assert sn_g == sn_o
        assert same_bits_f32(got.view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("intype,outtype", TYPE_PAIRS)
def test_block_schedule_track_mode(oracle, mixer, intype, outtype):
    """One shift per 8192-byte block, samplenum carried across shift changes (main.rs:177)."""
    rng = np.random.default_rng(21)
    fs = 1_024_000
    shifts = np.concatenate([np.repeat(np.float32(-9876.54), 60), np.repeat(np.float32(-9871.02), 55),
                             rng.uniform(-12000, 12000, 17).astype(np.float32), np.repeat(np.float32(5000.0), 40)])
    nbytes = shifts.size * BUFFER_SIZE - BPS[intype] * 333  # short last block
    buf = make_input(rng, nbytes // BPS[intype], intype)
    got, sn = mixer.mix_blocks(buf, intype, outtype, shifts, fs)
    want, sn_ref = oracle.mix_blocks(buf, intype, outtype, shifts, fs)
    assert sn == sn_ref
    check(oracle, got, want, outtype)


def test_track_replay_overpass(oracle, mixer):
    """cfg3 (cut): analytic overpass Doppler table -> the reference's replay driver (oracle) gives
    bytes + the per-block shift schedule; the planned entry point must reproduce the bytes."""
    fs = 1_024_000
    secs = 6
    t = np.arange(secs + 2, dtype=np.float64)
    v, d, tc, ftx = 7500.0, 700e3, 3.0, 437_505_000.0
    rr_km_s = v * v * (t - tc) / np.sqrt(d * d + (v * (t - tc)) ** 2) / 1000.0
    table = np.array([oracle.doppler_hz(x, 437_505_000) for x in rr_km_s])
    rng = np.random.default_rng(1024000)
    n = secs * fs
    tt = np.arange(n)
    sig = 0.25 * np.exp(2j * np.pi * 15000.0 / fs * tt) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty(2 * n, dtype=np.int16)
    iq[0::2] = np.clip(np.round(sig.real * 32767), -32768, 32767)
    iq[1::2] = np.clip(np.round(sig.imag * 32767), -32768, 32767)
    want, sn_ref, shifts, panicked = oracle.track_replay_stream(iq.view(np.uint8), I16, I16, table, 5000, fs)
    assert not panicked
    got, sn = mixer.mix_blocks(iq.view(np.uint8), I16, I16, shifts, fs)
    assert sn == sn_ref
    check(oracle, got, want, I16)


def test_large_period_l2_table_and_direct_paths_agree(oracle, mixer):
    """Same stream through (a) one call long enough to build a table and (b) many short calls that
    evaluate sincosf directly must give identical bytes (tables are built by the same routine)."""
    rng = np.random.default_rng(8)
    n = 600_000
    buf = make_input(rng, n, F32)
    shift, fs = -9876.54, 1_024_000  # period 111 145
    whole, sn_w = mixer.mix(buf, F32, F32, shift, fs)
    m2 = doppler_b200.Mixer(0)  # fresh context: no cached table
    sn = 0
    parts = []
    for k in range(0, n, 50_000):
        out, sn = m2.mix(buf[8 * k:8 * (k + 50_000)], F32, F32, shift, fs, samplenum=sn)
        parts.append(out)
    m2.close()
    assert sn == sn_w
    assert same_bits_f32(np.concatenate(parts), whole)
    want, _ = oracle.mix(buf, F32, F32, shift, fs)
    check(oracle, whole, want, F32)


def test_device_sincosf_matches_host_libm(oracle, mixer):
    """The kernel's sincosf against this box's libm over a strided sweep of all float bit patterns
    plus dense windows at the branch points."""
    for first, stride, count in [(0, 4099, 2**32 // 4099), (0x3F490FDB - 50_000, 1, 100_000), (0x42F00000 - 50_000, 1, 100_000),
                                 (0xBF490FDB - 50_000, 1, 100_000), (0xC2F00000 - 50_000, 1, 100_000), (0x7F7F0000, 1, 70_000),
                                 (0x39800000 - 50_000, 1, 100_000)]:
        s, c = mixer.sincosf_probe(first, stride, count)
        bits = (np.uint64(first) + np.arange(count, dtype=np.uint64) * np.uint64(stride)).astype(np.uint32)
        s_ref, c_ref = oracle.sincosf_batch(bits.view(np.float32))
        assert same_bits_f32(s, s_ref) and same_bits_f32(c, c_ref)


@pytest.mark.parametrize("shift,fs,n0,count", [(-15000.0, 256000, 0, 1000), (7321.7, 1_024_000, 0, 120_000),
                                               (4_000_000.5, 200_000_000, 1, 1_000_000), (-9876.54, 1_024_000, 2**24 - 1000, 5000),
                                               (1.0, 2_000_000_000, 2**32 - 2000, 4000)])
def test_device_phasor_matches_reference_ccexpf(oracle, mixer, shift, fs, n0, count):
    """(cos, sin) of theta_f32(n) as the reference forms it (dsp.rs:121-122)."""
    r = np.float32(shift) / np.float32(fs)
    c, s = mixer.phasor_probe(float(r), n0, count)
    n = (np.uint64(n0) + np.arange(count, dtype=np.uint64)).astype(np.uint32)
    theta = (np.float32(-2.0) * np.float32(np.pi)) * (r * n.astype(np.float32))
    s_ref, c_ref = oracle.sincosf_batch(theta)
    assert same_bits_f32(s, s_ref) and same_bits_f32(c, c_ref)
    re, im = oracle.ccexpf(0.0, float(theta[count // 2]))
    assert np.float32(re).tobytes() == c[count // 2].tobytes() and np.float32(im).tobytes() == s[count // 2].tobytes()


# (shift, fs, samplenum at the start, samples): table-less pieces, chosen so that whole tiles fall in each
# glibc range of the direct-evaluation fast rows (mixer_kernels.cuh): LARGE with theta < 0 and > 0 and
# several binade crossings, MEDIUM, SMALL, TINY, the f32(n) plateau above 2^24, the u32 wrap of
# samplenum, a period wrap inside a table-less periodic piece, and |r| > 1.
DIRECT_CASES = [
    (4_000_000.5, 200_000_000, 0, 3_000_000),        # P = 4.9 M > samples: tiny -> small -> medium -> large
    (-3_912_345.25, 200_000_000, 0, 50_000),         # P = 26 787, too short for a table: period wraps
    (7321.7, 1_024_000, 0, 100_000),                 # P = 55 244, 1.8 periods: no table, theta < 0
    (-9876.54, 1_024_000, 0, 200_000),               # P = 111 145: theta > 0, large range
    (1.0, 2_000_000_000, 0, 2_000_000),              # never resets: tiny then small
    (1.0, 2_000_000_000, 2**24 - 70_000, 300_000),   # f32(n) plateau
    (3.0, 2_000_000_000, 2**32 - 100_000, 250_000),  # samplenum wraps to 0 (release-mode u32)
    (1000.25, 48_000, 0, 150_000),                   # medium -> large within a few hundred samples
    (123_456.7, 48_000, 5, 60_000),                  # |r| > 1: large from the second sample on
    (-0.37, 1_000_000, 1_000_000, 400_000),          # small/medium boundary region
    # plateaus of f32(samplenum) above 2^24 at launch sizes that take the bulk-async direct shape: one phasor per distinct
    # f32(n) of a tile, parked in shared memory (mixer_kernels.cuh, stream_tile_direct)
    (1.0, 2_000_000_000, 2**24 - 1000, 5_000_000),   # crossing into the plateau region (pairs)
    (1.0, 2_000_000_000, 2**26 + 12_345, 5_000_000), # plateaus of 8, binade crossing at 2^26 + 2^26
    (3.0, 2_000_000_000, 2**31 - 2_000_000, 4_500_000),   # 128 -> 256 wide plateaus across 2^31
    (1.0e-3, 4_000_000_000, 2**32 - 3_000_000, 5_000_000),   # up to f32(n) = 2^32, the u32 wrap, the reset at 0, then tiny angles
]


@pytest.mark.parametrize("shift,fs,sn0,n", DIRECT_CASES)
def test_direct_fast_rows_match_oracle(oracle, mixer, shift, fs, sn0, n):
    """Unit samples (1 + 0j) make the f32 output the phasor itself, so every bit of every sincosf result
    of the direct path is compared with the reference's ccexpf; then random data through all type pairs."""
    ones = np.zeros(2 * n, dtype=np.float32)
    ones[0::2] = 1.0
    got, sn = mixer.mix(ones.view(np.uint8), F32, F32, shift, fs, samplenum=sn0)
    want, sn_ref = oracle.mix(ones.view(np.uint8), F32, F32, shift, fs, samplenum=sn0)
    assert sn == sn_ref
    check(oracle, got, want, F32)
    rng = np.random.default_rng(n)
    m = min(n, 300_000)
    for intype, outtype in TYPE_PAIRS:
        buf = make_input(rng, m, intype)
        got, sn = mixer.mix(buf, intype, outtype, shift, fs, samplenum=sn0)
        want, sn_ref = oracle.mix(buf, intype, outtype, shift, fs, samplenum=sn0)
        assert sn == sn_ref
        check(oracle, got, want, outtype)


# Periods above the shared-memory table size with several whole periods in the launch: COLUMN segments
# (phasors of a column evaluated once, parked in shared memory, reused over rows; mixer_kernels.cuh).
COLUMN_CASES = [   # launches of at least 4 Mi samples (smaller ones are latency-bound and stay on the GRID path)
    (-9876.54, 1_024_000, 0, 5_500_003),            # P = 111 145 (odd): rows at every alignment shift
    (7321.7, 1_024_000, 17, 4_300_001),             # P = 55 244, starts mid-period
    (-3_912_345.25, 200_000_000, 0, 4_700_000),     # P = 26 787
    (12_345.678, 1_024_000, 0, 4_200_000),
    (-1234.5, 96_000, 3, 4_900_002),
]


@pytest.mark.parametrize("shift,fs,sn0,n", COLUMN_CASES)
def test_column_segments_match_oracle(oracle, mixer, shift, fs, sn0, n):
    from doppler_b200 import dsp
    _, _, _, stats = dsp.plan_tiles_trace(F32, F32, sn0, [shift], n, fs, n)
    assert stats["column_segments"] >= 1, stats   # the case must actually take the COLUMN path
    ones = np.zeros(2 * n, dtype=np.float32)
    ones[0::2] = 1.0
    got, sn = mixer.mix(ones.view(np.uint8), F32, F32, shift, fs, samplenum=sn0)
    want, sn_ref = oracle.mix(ones.view(np.uint8), F32, F32, shift, fs, samplenum=sn0)
    assert sn == sn_ref
    check(oracle, got, want, F32)
    rng = np.random.default_rng(n)
    for intype, outtype in TYPE_PAIRS:
        buf = make_input(rng, n, intype)
        got, sn = mixer.mix(buf, intype, outtype, shift, fs, samplenum=sn0)
        want, sn_ref = oracle.mix(buf, intype, outtype, shift, fs, samplenum=sn0)
        assert sn == sn_ref
        check(oracle, got, want, outtype)


def test_column_segments_in_a_track_schedule(oracle, mixer):
    """Several long-period pieces in one launch (one shift per second at 1.024 Msps, as the replay driver
    produces), all four type pairs."""
    rng = np.random.default_rng(77)
    fs = 1_024_000
    per_sec = fs * 4 // BUFFER_SIZE   # i16 blocks per second
    shifts = np.concatenate([np.repeat(np.float32(s), per_sec) for s in (-9876.54, -9871.02, 7321.7, 5000.0, -3211.11)])
    for intype, outtype in TYPE_PAIRS:
        nbytes = shifts.size * BUFFER_SIZE - BPS[intype] * 777
        if intype == F32:
            nbytes = shifts.size * BUFFER_SIZE - 8 * 333   # f32 blocks hold half as many samples: 2.5 s of stream (GRID path)
        buf = make_input(rng, nbytes // BPS[intype], intype)
        got, sn = mixer.mix_blocks(buf, intype, outtype, shifts, fs)
        want, sn_ref = oracle.mix_blocks(buf, intype, outtype, shifts, fs)
        assert sn == sn_ref
        check(oracle, got, want, outtype)


def test_random_schedules_all_type_pairs(oracle, mixer):
    """Randomised per-block schedules (short / long / no reset periods mixed in one launch, arbitrary start
    samplenum, ragged ends) through the planned entry point, all four type pairs."""
    rng = np.random.default_rng(20151123)
    pool = np.array([-9876.54, 7321.7, 5000.0, -3211.11, -15000.0, 0.0, 12_345.678, 1.0, 815000.0, -1234.5], dtype=np.float32)
    for trial in range(12):
        intype, outtype = TYPE_PAIRS[trial % 4]
        fs = int(rng.choice([96_000, 1_024_000, 2_400_000]))
        nruns = int(rng.integers(1, 6))
        shifts = np.concatenate([np.repeat(rng.choice(pool), int(rng.integers(1, 1400))) for _ in range(nruns)])   # up to ~14 M samples
        nbytes = shifts.size * BUFFER_SIZE - BPS[intype] * int(rng.integers(0, BUFFER_SIZE // BPS[intype]))
        start = int(rng.choice([0, 1, 77_777, 2**24 + 5, 2**32 - 3]))
        buf = make_input(rng, nbytes // BPS[intype], intype)
        got, sn = mixer.mix_blocks(buf, intype, outtype, shifts, fs, samplenum=start)
        want, sn_ref = oracle.mix_blocks(buf, intype, outtype, shifts, fs, samplenum=start)
        assert sn == sn_ref, trial
        check(oracle, got, want, outtype)


def test_host_path_multi_chunk_with_long_periods(oracle, mixer):
    """Host-buffer entry point over several 32 MiB pipeline chunks: every chunk is planned on its own (pieces
    clipped at the chunk, COLUMN segments rebuilt, samplenum carried), the bytes must not show the seams."""
    rng = np.random.default_rng(31)
    n = 19_000_003                       # 76 MB of i16 input: three chunks, ragged end
    buf = make_input(rng, n, I16)
    for shift, fs in [(-9876.54, 1_024_000), (-15000.0, 256000)]:
        got, sn = mixer.mix(buf, I16, I16, shift, fs, samplenum=3)
        want, sn_ref = oracle.mix(buf, I16, I16, shift, fs, samplenum=3)
        assert sn == sn_ref
        check(oracle, got, want, I16)
    per_sec = 1_024_000 * 4 // BUFFER_SIZE
    shifts = np.concatenate([np.repeat(np.float32(s), per_sec) for s in np.linspace(-9000.0, 9000.0, 19)])
    nbytes = min(buf.size, shifts.size * BUFFER_SIZE - 4 * 123)
    got, sn = mixer.mix_blocks(buf[:nbytes], I16, F32, shifts, 1_024_000)
    want, sn_ref = oracle.mix_blocks(buf[:nbytes], I16, F32, shifts, 1_024_000)
    assert sn == sn_ref
    check(oracle, got, want, F32)


def test_plateau_piece_inside_a_segmented_launch(oracle, mixer):
    """A launch with COLUMN segments whose first piece is a never-resetting run above 2^24: the segmented kernel takes
    the plateau path for its GRID tiles, with the COLUMN window as scratch."""
    rng = np.random.default_rng(77)
    fs = 1_024_000
    bs = BUFFER_SIZE // 4
    shifts = np.concatenate([np.repeat(np.float32(1.0e-4), 1000), np.repeat(np.float32(-9876.54), 2200)])   # ~2 M + ~4.5 M samples
    n = shifts.size * bs - 333
    for intype, outtype in [(I16, I16), (F32, F32)]:
        bsz = BUFFER_SIZE // BPS[intype]
        m = min(n, shifts.size * bsz - 333)
        buf = make_input(rng, m, intype)
        got, sn = mixer.mix_blocks(buf, intype, outtype, shifts, fs, samplenum=2**25 + 7)
        want, sn_ref = oracle.mix_blocks_threads(buf, intype, outtype, shifts, fs, samplenum=2**25 + 7)
        assert sn == sn_ref
        check(oracle, got, want, outtype)


@pytest.mark.parametrize("shift,fs", [(float("inf"), 48000), (float("-inf"), 48000), (float("nan"), 48000), (1000.0, 0), (0.0, 0),
                                      (3.0e38, 1), (1.0e-30, 4_000_000_000)])
def test_degenerate_ratios(oracle, mixer, shift, fs):
    """r = shift / fs that is Inf, NaN (x / 0, Inf / fs), huge or denormal-small: the reference just runs its f32
    arithmetic (NaN phasors -> NaN f32 output, 0 after the saturating i16 cast); so must the kernels."""
    rng = np.random.default_rng(3)
    for intype, outtype in TYPE_PAIRS:
        for n in (5, 70_001):
            buf = make_input(rng, n, intype)
            got, sn = mixer.mix(buf, intype, outtype, shift, fs)
            want, sn_ref = oracle.mix(buf, intype, outtype, shift, fs)
            assert sn == sn_ref
            check(oracle, got, want, outtype)


@pytest.mark.parametrize("intype,outtype", TYPE_PAIRS)
def test_small_kernel_and_zero_copy_path_match_oracle(oracle, intype, outtype):
    """The latency-shaped small kernel (device launches up to SMALL_MAX_SAMPLES) and the zero-copy per-block host path
    (calls up to TINY_HOST_BYTES), forced on for every size here, including pieces that straddle groups, ragged ends,
    pinned caller buffers at odd offsets and a per-block schedule."""
    import ctypes
    from doppler_b200 import _lib
    lib = _lib.load()
    m = doppler_b200.Mixer(0)
    m.tune(small_max_samples=1 << 30, tiny_host_bytes=8 << 20)
    rng = np.random.default_rng(31)
    try:
        for shift, fs in SHIFTS[:5] + [(1.0, 2_000_000_000)]:
            for n in (1, 2, 3, 5, 2048, 2049, 30_001, 262_147):
                buf = make_input(rng, n, intype)
                got, sn = m.mix(buf, intype, outtype, shift, fs, samplenum=11)
                want, sn_ref = oracle.mix(buf, intype, outtype, shift, fs, samplenum=11)
                assert sn == sn_ref
                check(oracle, got, want, outtype)
        # one shift per 8192-byte block inside one tiny call
        bs = BUFFER_SIZE // BPS[intype]
        n = 9 * bs + 17
        shifts = rng.uniform(-12000, 12000, 10).astype(np.float32)
        buf = make_input(rng, n, intype)
        got, sn = m.mix_blocks(buf, intype, outtype, shifts, 1_024_000)
        want, sn_ref = oracle.mix_blocks(buf, intype, outtype, shifts, 1_024_000)
        assert sn == sn_ref
        check(oracle, got, want, outtype)
        # pinned caller buffers: used in place when 16-byte aligned, staged when not
        nb = 4096 * BPS[intype]
        hin, hout = lib.doppler_b200_host_alloc(nb + 64), lib.doppler_b200_host_alloc(2 * nb + 64)
        for off in (0, 8, 16):
            src = np.ctypeslib.as_array(ctypes.cast(hin + off, ctypes.POINTER(ctypes.c_uint8)), shape=(nb,))
            src[:] = make_input(rng, 4096, intype)
            ob = 4096 * BPS[outtype]
            dst = np.ctypeslib.as_array(ctypes.cast(hout + off, ctypes.POINTER(ctypes.c_uint8)), shape=(ob,))
            snc, gotc = ctypes.c_uint32(0), ctypes.c_size_t(0)
            rc = lib.doppler_b200_mix(m._ctx, hin + off, nb, intype, outtype, ctypes.c_float(7321.7), 1_024_000, ctypes.byref(snc), hout + off, ob,
                                      ctypes.byref(gotc))
            assert rc == 0 and gotc.value == ob
            want, sn_ref = oracle.mix(src.copy(), intype, outtype, 7321.7, 1_024_000)
            assert snc.value == sn_ref
            check(oracle, dst.copy(), want, outtype)
        lib.doppler_b200_host_free(hin)
        lib.doppler_b200_host_free(hout)
    finally:
        m.close()
