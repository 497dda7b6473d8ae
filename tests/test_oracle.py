"""Pins the CPU oracle (oracle/doppler_oracle.c) before anything is compared against it.

* every known answer the reference owns for this path: test_cexpf, /root/reference/src/dsp.rs:57-83
  (relative tolerance 1e-6 there; here also the exact bit patterns produced by the reference's
  complex.c compiled unmodified on glibc 2.39, recorded in SURVEY.md section 8c);
* the restated ccexpf against the reference's compiled complex.c (oracle/_ref) on a sweep;
* the Rust-semantics pieces that have no reference test: converters (dsp.rs:85-115), the
  samplenum reset rule (dsp.rs:125-130), saturating egress (main.rs:73-87), 8192-byte framing
  and the stop rule (main.rs:62-119), replay timing (main.rs:155-184).
"""
import math
import struct

import numpy as np
import pytest

from tests.oracle_lib import BUFFER_SIZE, F32, I16


def bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


# dsp.rs:57-83 -- (input re, im) -> (expected re, im), tolerance 1e-6 relative
TEST_CEXPF = [
    ((0.0, 0.0), (1.0, 0.0)),
    ((1.0, 1.0), (1.468694, 2.2873552)),
    ((70.0, 70.0), (1593075600000000000000000000000.0, 1946674600000000000000000000000.0)),
]
# exact bit patterns from the reference's own complex.c on this image's glibc
TEST_CEXPF_BITS = [
    ((0.0, 0.0), (0x3F800000, 0x00000000)),
    ((1.0, 1.0), (0x3FBBFE29, 0x40126407)),
    ((70.0, 70.0), (0x71A0DC0A, 0x71C4905C)),
    ((0.0, -0.5), (0x3F60A940, 0xBEF57744)),
    ((0.0, -100000.125), (0xBF7EFB32, 0x3DB68742)),
]


@pytest.mark.parametrize("which", ["restated", "reference"])
def test_cexpf_known_answers(oracle, which):
    if which == "reference" and not oracle.has_ref:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    f = oracle.ccexpf_restated if which == "restated" else oracle.ccexpf_reference
    for (a, b), (er, ei) in TEST_CEXPF:
        re, im = f(a, b)
        if er == 0.0:
            assert re == er
        else:
            assert abs((re - er) / er) < 1e-6  # assert_eq_delta, dsp.rs:50-55
        if ei == 0.0:
            assert im == ei
        else:
            assert abs((im - ei) / ei) < 1e-6
    re, im = f(1_000_000.0, 1_000_000.0)  # dsp.rs:77-80
    assert re == math.inf and im == -math.inf
    for (a, b), (br, bi) in TEST_CEXPF_BITS:
        re, im = f(a, b)
        assert (bits(re), bits(im)) == (br, bi), (a, b, hex(bits(re)), hex(bits(im)))


def test_restated_ccexpf_equals_reference_object_code(oracle):
    if not oracle.has_ref:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1)
    thetas = np.concatenate([
        rng.uniform(-7, 7, 2000), rng.uniform(-130, 130, 2000), rng.uniform(-7e5, 7e5, 4000),
        np.array([0.0, -0.0, 1e-45, -1e-38, 1e-13, 0.785398, 119.99999, 120.0, 3.4e38, np.inf, -np.inf, np.nan]),
    ]).astype(np.float32)
    for t in thetas:
        a = oracle.ccexpf_restated(0.0, float(t))
        b = oracle.ccexpf_reference(0.0, float(t))
        if math.isnan(b[0]) or math.isnan(b[1]):
            assert math.isnan(a[0]) == math.isnan(b[0]) and math.isnan(a[1]) == math.isnan(b[1])
        else:
            assert (bits(a[0]), bits(a[1])) == (bits(b[0]), bits(b[1])), t


def test_ccexpf_zero_real_is_sincosf(oracle):
    """cexpf(0 + i*theta) == (cosf, sinf)(theta) bit for bit -- the identity the GPU kernel relies on."""
    rng = np.random.default_rng(2)
    th = np.concatenate([rng.uniform(-200, 200, 3000), rng.uniform(-1e6, 1e6, 3000), [0.0, -0.0, 1e-40, -3e-39]]).astype(np.float32)
    s, c = oracle.sincosf_batch(th)
    for t, si, ci in zip(th, s, c):
        re, im = oracle.ccexpf(0.0, float(t))
        assert (bits(re), bits(im)) == (bits(float(ci)), bits(float(si))), t


def test_convert_i16(oracle):
    raw = np.array([0, 1, -1, 32767, -32768, 256, -256, 12345], dtype="<i2")
    out = oracle.convert_iqi16_to_complex(raw.tobytes())
    want = (raw.astype(np.float32) / np.float32(32768.0)).view(np.complex64)
    assert np.array_equal(out, want)
    assert oracle.convert_iqi16_to_complex(b"\x00" * 6) is None  # assert!(len % 4 == 0), dsp.rs:87
    assert oracle.convert_iqi16_to_complex(b"").size == 0


def test_convert_f32_is_bit_preserving(oracle):
    pat = np.array([0x00000000, 0x80000000, 0x7FC00001, 0xFFC12345, 0x7F800000, 0x00000001, 0x3F800000, 0xC2F6E979], dtype="<u4")
    out = oracle.convert_iqf32_to_complex(pat.tobytes())
    assert np.array_equal(out.view(np.uint32), pat)
    assert oracle.convert_iqf32_to_complex(b"\x00" * 12) is None  # assert!(len % 8 == 0), dsp.rs:103


def test_samplenum_rule_known_periods(oracle):
    """Reset periods measured on the reference arithmetic."""
    for shift, fs, period in [(-15000.0, 256000, 256), (100000.0, 10_000_000, 100), (815000.0, 2_400_000, 480), (5000.0, 1_024_000, 1024)]:
        tr, sn = oracle.samplenum_trace(0, shift, fs, 3 * period + 5)
        assert tr[0] == 0 and tr[1] == 1
        assert np.array_equal(tr[1:1 + period], np.arange(1, period + 1))
        assert tr[1 + period] == 1
        assert sn == oracle.samplenum_advance(0, shift, fs, 3 * period + 5)


def test_shift_frequency_matches_scalar_formula(oracle):
    rng = np.random.default_rng(3)
    x = (rng.uniform(-1, 1, 64) + 1j * rng.uniform(-1, 1, 64)).astype(np.complex64)
    out, sn = oracle.shift_frequency(x, 0, -15000.0, 256000)
    assert sn == 64
    r = np.float32(-15000.0) / np.float32(256000)
    for k in range(64):
        n = 0 if k == 0 else k
        th = np.float32(-2.0) * np.float32(np.pi) * (r * np.float32(n))
        s, c = oracle.sincosf_batch(np.array([th], dtype=np.float32))
        a, b = np.float32(x[k].real), np.float32(x[k].imag)
        re = np.float32(a * c[0]) - np.float32(b * s[0])
        im = np.float32(a * s[0]) + np.float32(b * c[0])
        assert out[k].real == re and out[k].imag == im


def test_egress_i16_saturates_like_rust_as(oracle):
    # (v * 32767.0) as i16: truncation toward zero, saturation, NaN -> 0 (main.rs:77-78)
    vals = np.array([0.0, 0.99999, -0.99999, 1.0, -1.0, 1.5, -1.5, np.nan, np.inf, -np.inf, 3.0517578e-05, -3.0517578e-05,
                     0.5000153, -1.00003], dtype=np.float32)
    def rust_as_i16(f):
        f = float(np.float32(f) * np.float32(32767.0))
        if math.isnan(f):
            return 0
        if math.isinf(f):
            return 32767 if f > 0 else -32768
        return int(max(-32768, min(32767, math.trunc(f))))

    # I channel: (v, 0) * (1, -0) -> re = v;  Q channel: (0, v) * (1, -0) -> im = v
    x = np.zeros(2 * vals.size, dtype=np.complex64)
    x.real[:vals.size] = vals
    x.imag[vals.size:] = vals
    out, _ = oracle.mix(x.view(np.uint8), F32, I16, 0.0, 48000)  # shift 0 -> multiply by (1, -0)
    got = out.view("<i2").reshape(-1, 2)
    for k, v in enumerate(vals):
        assert got[k, 0] == rust_as_i16(v), (v, got[k, 0])
        assert got[vals.size + k, 1] == rust_as_i16(v), (v, got[vals.size + k, 1])


@pytest.mark.parametrize("nsamples", [256_000, 255_999, 2048, 2047, 1, 0])
def test_const_stream_framing(oracle, nsamples):
    """main.rs:62-119: 8192-byte blocks, stop on the first short read; the block pump must give
    the same bytes as one fused call over the whole buffer (shift is constant, samplenum carried)."""
    rng = np.random.default_rng(nsamples + 7)
    iq = rng.integers(-20000, 20000, 2 * nsamples, dtype=np.int16)
    out, sn, panicked = oracle.const_stream(iq.view(np.uint8), I16, I16, -15000, 256000)
    assert not panicked
    whole, sn2 = oracle.mix(iq.view(np.uint8), I16, I16, -15000.0, 256000)
    assert np.array_equal(out, whole) and sn == sn2
    assert out.size == 4 * nsamples


def test_const_stream_misaligned_tail_panics_after_flushing_full_blocks(oracle):
    buf = np.zeros(BUFFER_SIZE + 6, dtype=np.uint8)  # f32 input, tail of 6 bytes
    out, sn, panicked = oracle.const_stream(buf, F32, F32, 1000, 48000)
    assert panicked and out.size == BUFFER_SIZE  # first block written and flushed (main.rs:97) before dsp.rs:103 fires


def test_track_replay_shift_schedule(oracle):
    """main.rs:155-184: block i uses the Doppler evaluated at the dt computed during block i-1;
    dt = trunc(f32(sample_count)/f32(fs)) whole seconds."""
    fs = 4096  # 2 blocks of i16 per second
    table = np.array([1000.25, 2000.5, 3000.75, 4000.0])
    nblocks = 9
    buf = np.zeros(nblocks * BUFFER_SIZE, dtype=np.uint8)
    out, sn, shifts, panicked = oracle.track_replay_stream(buf, I16, I16, table, 5, fs)
    assert not panicked
    # blocks 0,1 -> dt 0 (block 1 still sees dt computed from sample_count 0); block 2 sees dt from count 2048 -> 0;
    # block 3 sees dt from count 4096 -> 1; ...  plus the final empty block
    exp_dt = [0, 0, 0, 1, 1, 2, 2, 3, 3, 3]
    want = [np.float32(table[d]) + np.float32(5) for d in exp_dt]
    assert len(shifts) == nblocks + 1
    assert [float(s) for s in shifts] == [float(w) for w in want]
    via_blocks, sn2 = oracle.mix_blocks(buf, I16, I16, shifts, fs)
    assert np.array_equal(via_blocks, out) and sn == sn2


def test_doppler_hz_formula(oracle):
    # main.rs:163 with the README's ESTCube-1 carrier
    assert oracle.doppler_hz(-7.0, 437505000) == (-7.0 * 1000.0 / 299792458.0) * 437505000.0 * (-1.0)
