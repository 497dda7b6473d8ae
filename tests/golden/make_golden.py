#!/usr/bin/env python
"""Regenerates tests/golden/*.{json,npz}.  Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

What is pinned and by what:
  * cexpf_known_answers.json -- the reference's ONLY known-answer test, test_cexpf
    (/root/reference/src/dsp.rs:57-83): inputs and expected values transcribed from the test,
    plus the exact bit patterns that the reference's own src/complex.c (compiled unmodified into
    oracle/_ref/libcomplex_ref.so) returns on this image's glibc.
  * mixer_vectors.npz -- the reference holds NO vector for the mixer loop, the converters, the
    samplenum rule or the egress casts.  These vectors are produced by the
    oracle's restatement of dsp.rs:85-134 + main.rs:62-119,155-184 with its per-sample ccexpf
    call bound to the reference's compiled complex.c, i.e. by executing the reference's own
    native code for the arithmetic core.  They freeze that behaviour so that the GPU box (no
    /root/reference there) and later rounds check against bytes, not against a rebuilt oracle.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.oracle_lib import BUFFER_SIZE, F32, I16, Oracle  # noqa: E402


def tone_i16(n, fs, seed, f0=15000.0):
    """SURVEY 8(d) cfg1 signal: tone at +15 kHz, 0.25 FS, Gaussian noise sigma 0.05 FS."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    sig = 0.25 * np.exp(2j * np.pi * f0 / fs * t) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty(2 * n, dtype="<i2")
    iq[0::2] = np.clip(np.round(sig.real * 32767), -32768, 32767)
    iq[1::2] = np.clip(np.round(sig.imag * 32767), -32768, 32767)
    return iq.view(np.uint8)


def bits(x):
    return f"{np.float32(x).view(np.uint32):08x}"


def main():
    if not os.path.exists("/root/reference/src/complex.c"):
        sys.exit("needs /root/reference (authoring container)")
    import __graft_entry__
    __graft_entry__.build()
    oracle = Oracle(use_ref=True)
    assert oracle.using_ref, "oracle/_ref/libcomplex_ref.so was not built"

    # ---- test_cexpf, dsp.rs:57-83 ------------------------------------------------------------
    ka = []
    for (re, im), (ere, eim), tol in [((0.0, 0.0), (1.0, 0.0), 1e-6), ((1.0, 1.0), (1.468694, 2.2873552), 1e-6),
                                      ((70.0, 70.0), (1.5930756e30, 1.9466746e30), 1e-6),
                                      ((1e6, 1e6), (float("inf"), float("-inf")), 0.0)]:
        gre, gim = oracle.ccexpf_reference(re, im)
        ka.append({"in": [re, im], "expect": [repr(ere), repr(eim)], "rel_tol": tol, "source": "dsp.rs:57-83",
                   "ref_bits": [bits(gre), bits(gim)]})
    # the argument shape the mixer actually uses: real part 0 (dsp.rs:121)
    for theta in (-0.5, -100000.125, 0.0, -6.2831855, 3.0e5, -5.9e5, 1.0e-5, -7.7e37):
        gre, gim = oracle.ccexpf_reference(0.0, theta)
        ka.append({"in": [0.0, theta], "source": "probe through reference complex.c, real part 0 (dsp.rs:121)",
                   "ref_bits": [bits(gre), bits(gim)]})
    with open(os.path.join(HERE, "cexpf_known_answers.json"), "w") as f:
        json.dump({"glibc": oracle.libc_version(), "reference": "cubehub/doppler @ 5f13df14 src/complex.c (compiled unmodified)",
                   "vectors": ka}, f, indent=1)

    # ---- mixer vectors -----------------------------------------------------------------------
    out = {}
    meta = []
    rng = np.random.default_rng(20161017)

    def add(name, kind, params, inp, got, samplenum, extra=None):
        out[name + ".in"] = np.frombuffer(bytes(inp), dtype=np.uint8).copy()
        out[name + ".out"] = np.asarray(got, dtype=np.uint8).copy()
        m = {"name": name, "kind": kind, "samplenum_out": samplenum}
        m.update(params)
        if extra:
            m.update(extra)
        meta.append(m)

    # cfg1 cut: const stream, i16 -> i16, three full blocks + a short one
    inp = tone_i16(3 * 2048 + 777, 256000, 20150122)
    got, sn, pan = oracle.const_stream(inp, I16, I16, -15000, 256000)
    assert not pan
    add("cfg1_const_i16_i16", "const_stream", {"intype": I16, "outtype": I16, "shift": -15000, "samplerate": 256000}, inp, got, sn)
    # exact multiple of the block: the reference ends on an empty read (main.rs:98)
    inp = tone_i16(2 * 2048, 256000, 5)
    got, sn, pan = oracle.const_stream(inp, I16, F32, -15000, 256000)
    add("const_exact_blocks_i16_f32", "const_stream", {"intype": I16, "outtype": F32, "shift": -15000, "samplerate": 256000}, inp, got, sn)
    # cfg2 cut: f32 -> i16 @ 10 Msps, shift 100000
    x = np.random.default_rng(10_000_000).uniform(-0.7, 0.7, 2 * (4 * 1024 + 333)).astype("<f4").view(np.uint8)
    got, sn, pan = oracle.const_stream(x, F32, I16, 100000, 10_000_000)
    add("cfg2_const_f32_i16", "const_stream", {"intype": F32, "outtype": I16, "shift": 100000, "samplerate": 10_000_000}, x, got, sn)
    # the reference bench's parameters (dsp.rs:136-157): 815 kHz @ 2.4 Msps, f32 0xAA bytes
    x = np.full(8 * 3000, 0xAA, dtype=np.uint8)
    got, sn = oracle.mix(x, F32, F32, 815000.0, 2_400_000)
    add("bench_params_f32_f32", "mix", {"intype": F32, "outtype": F32, "shift_hz": 815000.0, "samplerate": 2_400_000, "samplenum_in": 0}, x, got, sn)
    # irregular ratio (large period, large-argument trig), all four type pairs
    for it, ot in [(I16, I16), (I16, F32), (F32, I16), (F32, F32)]:
        if it == I16:
            x = rng.integers(-32768, 32768, 2 * 5000, dtype=np.int32).astype("<i2").view(np.uint8)
        else:
            x = rng.uniform(-1.0, 1.0, 2 * 5000).astype("<f4").view(np.uint8)
        got, sn = oracle.mix(x, it, ot, 7321.7, 1_024_000, samplenum=50_000)
        add(f"irregular_{'i16' if it == I16 else 'f32'}_{'i16' if ot == I16 else 'f32'}", "mix",
            {"intype": it, "outtype": ot, "shift_hz": 7321.7, "samplerate": 1_024_000, "samplenum_in": 50_000}, x, got, sn)
    # f32 specials pass through the bit-copy ingest (dsp.rs:101-115)
    pat = np.array([0x7FC00000, 0x3F800000, 0x7F800000, 0x00000000, 0xFF800000, 0x3F000000, 0x00000001, 0x80000001,
                    0x00800000, 0x7F7FFFFF, 0x3F800000, 0x7F7FFFFF], dtype="<u4")
    x = np.tile(pat, 200).view(np.uint8)
    for ot in (I16, F32):
        got, sn = oracle.mix(x, F32, ot, 100000.0, 10_000_000)
        add(f"specials_f32_{'i16' if ot == I16 else 'f32'}", "mix",
            {"intype": F32, "outtype": ot, "shift_hz": 100000.0, "samplerate": 10_000_000, "samplenum_in": 0, "nan_payload_free": True}, x, got, sn)
    # full-scale saturation of the i16 egress (main.rs:77-78, Rust saturating `as`)
    x = rng.choice(np.array([-32768, -32767, 32767, 23170, -23170, 0, 1, -1], dtype="<i2"), 2 * 4000).view(np.uint8)
    got, sn = oracle.mix(x, I16, I16, 815000.0, 2_400_000)
    add("fullscale_i16_i16", "mix", {"intype": I16, "outtype": I16, "shift_hz": 815000.0, "samplerate": 2_400_000, "samplenum_in": 0}, x, got, sn)
    # samplenum handed in near the u32 wrap, tiny ratio (no reset below 2^26)
    x = rng.uniform(-1.0, 1.0, 2 * 3000).astype("<f4").view(np.uint8)
    got, sn = oracle.mix(x, F32, F32, 1.0, 2_000_000_000, samplenum=2**32 - 1500)
    add("wrap_f32_f32", "mix", {"intype": F32, "outtype": F32, "shift_hz": 1.0, "samplerate": 2_000_000_000, "samplenum_in": 2**32 - 1500}, x, got, sn)
    # track mode: one shift per 8192-byte block, samplenum carried across shift changes (main.rs:177)
    shifts = np.array([-9876.54, -9876.54, -9871.02, 5000.0, 0.0, 7321.7, 7321.7], dtype=np.float32)
    x = tone_i16(shifts.size * 2048 - 100, 1_024_000, 21)
    got, sn = oracle.mix_blocks(x, I16, I16, shifts, 1_024_000)
    add("track_blocks_i16_i16", "mix_blocks", {"intype": I16, "outtype": I16, "samplerate": 1_024_000, "samplenum_in": 0,
                                                 "shifts_hz_bits": [int(v) for v in shifts.view(np.uint32)]}, x, got, sn)
    # track replay driver (main.rs:155-184) over an analytic overpass table; fs chosen so the
    # whole-second clock (main.rs:166) ticks every 4 blocks
    fs = 8192
    secs = 5
    t = np.arange(secs + 2, dtype=np.float64)
    v, d, tc = 7500.0, 700e3, 2.0
    rr = v * v * (t - tc) / np.sqrt(d * d + (v * (t - tc)) ** 2) / 1000.0
    table = np.array([oracle.doppler_hz(r, 437_505_000) for r in rr])
    x = tone_i16(secs * fs - 300, fs, 1024000, f0=500.0)
    got, sn, used, pan = oracle.track_replay_stream(x, I16, F32, table, 5000, fs)
    assert not pan
    add("track_replay_i16_f32", "track_replay", {"intype": I16, "outtype": F32, "samplerate": fs, "offset": 5000,
                                                   "doppler_hz_by_second": [float(v) for v in table],
                                                   "range_rate_km_s": [float(v) for v in rr], "frequency": 437_505_000,
                                                   "shifts_hz_bits": [int(v) for v in used.view(np.uint32)]}, x, got, sn)

    np.savez_compressed(os.path.join(HERE, "mixer_vectors.npz"), **out)
    with open(os.path.join(HERE, "mixer_vectors.json"), "w") as f:
        json.dump({"glibc": oracle.libc_version(), "ccexpf": "reference src/complex.c compiled unmodified (oracle/_ref)",
                   "generator": "tests/golden/make_golden.py", "cases": meta}, f, indent=1)
    print(f"wrote {len(meta)} mixer cases, {len(ka)} cexpf vectors")


if __name__ == "__main__":
    main()
