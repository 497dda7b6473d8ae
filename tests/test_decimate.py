"""The fused downstream stage: mix + decimate-by-M FIR (include/doppler_b200.h: doppler_b200_decim_*).
Not in the reference: the oracle's oracle_mix_decimate is the specification, pinned on the CPU against an independent
double-precision convolution of the oracle's own mixer output; the CUDA path must reproduce it bit for bit."""
import numpy as np
import pytest

import doppler_b200
from doppler_b200 import F32, I16
from tests.oracle_lib import BUFFER_SIZE, same_bits_f32

BPS = {I16: 4, F32: 8}


def lowpass(ntaps, cutoff):
    t = np.arange(ntaps) - (ntaps - 1) / 2.0
    h = np.sinc(2 * cutoff * t) * np.hamming(ntaps)
    return (h / h.sum()).astype(np.float32)


def make_input(rng, n, typ):
    if typ == I16:
        return rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype("<i2").view(np.uint8)
    return rng.uniform(-0.7, 0.7, 2 * n).astype("<f4").view(np.uint8)


# ---- CPU: the specification itself -----------------------------------------------------------------

@pytest.mark.parametrize("M,ntaps", [(1, 1), (2, 5), (8, 33), (3, 16), (10, 4)])
def test_oracle_decimator_is_the_fir_of_the_oracle_mixer(oracle, M, ntaps):
    rng = np.random.default_rng(M * 100 + ntaps)
    n = 20_000 + M - 1
    x = make_input(rng, n, F32)
    taps = rng.uniform(-0.2, 0.2, ntaps).astype(np.float32)
    z, st = oracle.mix_decimate(x, F32, F32, 7321.7, 1_024_000, taps, M)
    y, sn = oracle.mix(x, F32, F32, 7321.7, 1_024_000)
    ref = np.convolve(y.view(np.complex64).astype(np.complex128), taps.astype(np.float64))[:n][::M]
    got = z.view(np.complex64)
    assert got.size == ref.size == (n + M - 1) // M
    assert np.abs(got - ref).max() < 1e-5                       # f32 accumulation against f64
    assert st["samplenum"] == sn and st["pos"] == n
    # chunked == whole, at cuts that are not multiples of M or of the tap count
    state, parts = None, []
    for a, b in ((0, 1), (1, 7), (7, 4096), (4096, 4099), (4099, n)):
        part, state = oracle.mix_decimate(x[8 * a:8 * b], F32, F32, 7321.7, 1_024_000, taps, M, state)
        parts.append(part)
    assert np.array_equal(np.concatenate(parts), z) and state["pos"] == n


def test_oracle_decimator_i16_egress_saturates_like_the_mixer(oracle):
    rng = np.random.default_rng(5)
    x = np.full(2 * 4096, 32767, dtype="<i2").view(np.uint8)
    taps = np.full(8, 0.5, dtype=np.float32)                     # gain 4: drives the i16 egress into saturation
    z, _ = oracle.mix_decimate(x, I16, I16, 0.0, 48000, taps, 4)
    v = z.view("<i2")
    assert v.max() == 32767 and v.size == 2 * 1024


# ---- GPU: parity with the specification ---------------------------------------------------------------

@pytest.fixture(params=["register-blocked kernel", "generic kernel"])
def decim_variant(request, mixer):
    """Both forms of the fused stage (doppler_b200_tune DECIM_VARIANT): the register-blocked kernel wherever the filter fits its
    envelope, and the generic kernel for every filter."""
    mixer.tune(decim_variant=int(request.param == "generic kernel"))
    yield request.param
    mixer.tune(decim_variant=0)


# (M, ntaps): every filter shape of the register-blocked kernel -- ntaps <= M, <= 2M, <= 3M, > 3M, ties at the multiples --
# odd and even M, M = 1, a stage of one warp (M = 64), and filters outside its envelope (M > 64, 3M + ntaps > 224)
FILTERS = [(8, 33), (4, 64), (3, 17), (1, 9), (25, 101), (2, 1), (8, 20), (8, 12), (8, 5), (8, 8), (8, 16), (8, 24), (5, 15),
           (7, 14), (16, 49), (64, 30), (6, 200), (10, 4), (100, 40), (2, 300)]


@pytest.mark.gpu
@pytest.mark.parametrize("intype,outtype", [(I16, I16), (I16, F32), (F32, I16), (F32, F32)])
@pytest.mark.parametrize("M,ntaps", FILTERS)
def test_mix_decimate_matches_oracle(oracle, mixer, decim_variant, intype, outtype, M, ntaps):
    rng = np.random.default_rng(M * 1000 + ntaps)
    taps = lowpass(ntaps, 0.4 / M) if ntaps > 1 else np.array([0.75], dtype=np.float32)
    dec = doppler_b200.Decimator(mixer, taps, M)
    try:
        for shift, fs in ((-15000.0, 256000), (7321.7, 1_024_000)):
            dec.reset()
            state, sn = None, 0
            for n in (1, M, 5, 2048, 70_001, 300_003):           # one stream, six calls: history and phase carried across them
                buf = make_input(rng, n, intype)
                got, sn = dec.mix(buf, intype, outtype, shift, fs, samplenum=sn)
                want, state = oracle.mix_decimate(buf, intype, outtype, shift, fs, taps, M, state)
                assert sn == state["samplenum"] and dec.position == state["pos"]
                assert got.size == want.size
                assert np.array_equal(got, want) if outtype == I16 else same_bits_f32(got, want), (shift, n)
    finally:
        dec.close()


@pytest.mark.gpu
def test_mix_blocks_decimate_and_chunked_host_pipeline(oracle, mixer, decim_variant):
    """A per-block shift schedule, and an input long enough for several 32 MiB pipeline chunks (history handed from slot to slot)."""
    rng = np.random.default_rng(77)
    M, taps = 8, lowpass(48, 0.05)
    dec = doppler_b200.Decimator(mixer, taps, M)
    try:
        fs, bs = 1_024_000, BUFFER_SIZE // 4
        shifts = np.repeat(rng.uniform(-12000, 12000, 12).astype(np.float32), 40)
        n = shifts.size * bs - 100
        buf = make_input(rng, n, I16)
        got, sn = dec.mix_blocks(buf, I16, F32, shifts, fs)
        want, st = oracle.mix_decimate(buf, I16, F32, shifts, fs, taps, M)
        assert sn == st["samplenum"] and same_bits_f32(got, want)
        dec.reset()
        n = 20_000_003                                              # 160 MB of f32: five chunks
        buf = make_input(rng, n, F32)
        got, sn = dec.mix(buf, F32, I16, 100000.0, 10_000_000)
        want, st = oracle.mix_decimate(buf, F32, I16, 100000.0, 10_000_000, taps, M)
        assert sn == st["samplenum"] and np.array_equal(got, want)
    finally:
        dec.close()


@pytest.mark.gpu
def test_mix_decimate_dev_and_argument_checks(oracle, mixer, decim_variant):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(9)
    M, taps = 5, lowpass(31, 0.08)
    dec = doppler_b200.Decimator(mixer, taps, M)
    try:
        n = 1_000_003
        buf = make_input(rng, n, F32)
        x = torch.from_numpy(buf.copy()).cuda()
        y = torch.empty((n // M + 2) * 8, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        nbytes, sn = dec.mix_dev(x.data_ptr(), x.numel(), F32, F32, -9876.54, 1_024_000, 0, y.data_ptr(), y.numel())
        mixer.synchronize()
        want, st = oracle.mix_decimate(buf, F32, F32, -9876.54, 1_024_000, taps, M)
        assert nbytes == want.size and sn == st["samplenum"]
        assert same_bits_f32(y[:nbytes].cpu().numpy(), want)
        with pytest.raises(doppler_b200.DopplerError) as ei:       # the reference's assert on ragged input (dsp.rs:103)
            dec.mix(np.zeros(13, dtype=np.uint8), F32, F32, 1.0, 48000)
        assert ei.value.code == doppler_b200.dsp.EALIGN
    finally:
        dec.close()
    with pytest.raises(doppler_b200.DopplerError):
        doppler_b200.Decimator(mixer, np.zeros(0, dtype=np.float32), 4)
    with pytest.raises(doppler_b200.DopplerError):
        doppler_b200.Decimator(mixer, taps, 0)


# ---- CPU: the register-blocked kernel's walk plan (host logic, no device) -----------------------------------------

def walk_trace(taps, M, first_out=0):
    import ctypes

    from doppler_b200 import _lib
    lib = _lib.load()
    h = np.ascontiguousarray(taps, dtype=np.float32)
    cap = 4 * M + h.size + 8
    rec = np.zeros((cap, 8), dtype=np.uint32)
    bits = np.zeros((cap, 4), dtype=np.uint32)
    info = np.zeros(4, dtype=np.uint32)
    n = lib.doppler_b200_decim_walk_trace(h.ctypes.data_as(ctypes.c_void_p), h.size, M, first_out, rec.ctypes.data_as(ctypes.c_void_p),
                                          bits.ctypes.data_as(ctypes.c_void_p), cap, info.ctypes.data_as(ctypes.c_void_p))
    return n, rec[:max(n, 0)], bits[:max(n, 0)], info


@pytest.mark.parametrize("M,ntaps", FILTERS + [(3, 1), (1, 1), (64, 1), (2, 218), (12, 13), (12, 24), (12, 25), (12, 36), (12, 37)])
@pytest.mark.parametrize("first_out", [0, 1, 5])
def test_walk_plan_is_the_fir_in_tap_order(M, ntaps, first_out):
    """The walk the kernel is launched with (tap layout in the kernel parameters, segment bounds, run list with the padding
    slots) replayed on the host: every output of a thread's group must meet taps 0 .. ntaps-1 in that order, each with the
    sample the specification pairs it with, at the shared-memory slot the staging phase wrote that sample to."""
    rng = np.random.default_rng(ntaps * 131 + M)
    taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
    n, rec, bits, info = walk_trace(taps, M, first_out % M)
    if M > 64 or 3 * M + ntaps > 224:
        assert n == 0                                             # outside the envelope: the generic kernel takes the launch
        return
    assert n == 3 * M + ntaps
    tb, lead, nt, shape = (int(v) for v in info)
    assert shape == min(3, (ntaps - 1) // M) and 32 <= tb <= nt and tb % 32 == 0 and nt in (128, 256)
    assert lead == (first_out % M - (ntaps - 1)) % 4              # the staging origin is a multiple of 4 samples
    RM = 4 * M
    c0 = lead + (ntaps - 1) + 3 * M
    seen = {k: [] for k in range(4)}
    for u in range(n):
        c, slot, klo, khi = (int(v) for v in rec[u, :4])
        assert c == c0 - u and slot == c + c // RM                # newest first; one padding slot per 4M staged samples
        for k in range(4):
            t = int(rec[u, 4 + k])
            active = klo <= k <= khi
            assert active == (0 <= u - (3 - k) * M < ntaps)
            if active:
                assert t == u - (3 - k) * M
                # output k of the group sits at staged index lead + (ntaps-1) + k*M; tap t pairs it with the sample t before
                assert c == lead + (ntaps - 1) + k * M - t
                assert bits[u, k] == taps[t:t + 1].view(np.uint32)[0]
                seen[k].append(t)
            else:
                assert t == 0xffffffff
    for k in range(4):
        assert seen[k] == list(range(ntaps))                      # the specification's summation order
    # the CTA's staged slots fit the shared-memory budget and threads' strides are odd in 8-byte units (conflict-free)
    count = lead + (4 * tb - 1) * M + ntaps
    assert count + count // RM + 2 <= 8448 and (RM + 1) % 2 == 1
