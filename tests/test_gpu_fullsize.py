"""BASELINE.json configs at their FULL sizes, device-resident, checked through properties that do not need
the oracle to mix the whole stream:

  * sampled windows: random windows (plus the first and the ragged last one) of the full-size output are
    compared bit for bit with the oracle run on the same input window, seeded with the analytic samplenum
    of the window's first sample (the same host function that seeds the time slices of a multi-GPU job);
  * chunked == whole: the second half of the stream mixed on its own, from its analytic seed, equals the
    second half of the one-call result byte for byte (torch.equal on the device).

Sizes: cfg2 = 640 M samples f32->i16 (the bench workload), cfg3 = 600 s @ 1.024 Msps i16 track replay
(614.4 M samples, ~600 long-period pieces in one launch), cfg4 = one GPU's slice of the 8-GPU job
(1.5 G samples f32->f32 @ 200 Msps, stream-absolute indices beyond 2^32, two launches)."""
import numpy as np
import pytest

import doppler_b200
from doppler_b200 import F32, I16, dsp, slicing
from tests.oracle_lib import BUFFER_SIZE

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

BPS = {I16: 4, F32: 8}
WINDOW = 65_536


def _windows(rng, n, align, count):
    """[begin, end) sample windows: the first, `count` random ones aligned to `align`, and the ragged last."""
    starts = [0] + [int(s) * align for s in rng.integers(1, (n - WINDOW) // align, count)]
    wins = [(s, min(s + WINDOW, n)) for s in starts]
    last = max(0, (n - WINDOW // 2) // align * align)
    wins.append((last, n))
    return wins


def _overpass_table(oracle, seconds, ftx, tc):
    t = np.arange(seconds + 2, dtype=np.float64)
    v, d = 7500.0, 700e3
    rr_km_s = v * v * (t - tc) / np.sqrt(d * d + (v * (t - tc)) ** 2) / 1000.0
    return np.array([oracle.doppler_hz(x, ftx) for x in rr_km_s])


def _fill(x, typ):
    if typ == F32:
        x.view(torch.float32).uniform_(-0.7, 0.7)
    else:
        v = x.view(torch.int16)
        step = 1 << 28
        for k in range(0, v.numel(), step):   # chunked: randint materialises int64 internally
            v[k:k + step].copy_(torch.randint(-20000, 20000, (min(step, v.numel() - k),), device=x.device, dtype=torch.int16))
    torch.cuda.synchronize()   # the mixer runs on its own non-blocking stream: the fill must have landed first


def _same(got, want, outtype):
    if outtype == I16:
        return np.array_equal(got, want)
    g, w = got.view(np.uint32), want.view(np.uint32)
    nan = np.isnan(got.view(np.float32)) & np.isnan(want.view(np.float32))
    return bool(np.all((g == w) | nan))


def test_cfg2_const_f32_to_i16_full_size(oracle, mixer):
    n, shift, fs = 640_000_003, 100000.0, 10_000_000   # the bench workload plus a ragged 3-sample end
    x = torch.empty(n * 8, dtype=torch.uint8, device="cuda")
    y = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
    _fill(x, F32)
    sn = mixer.mix_dev(x.data_ptr(), x.numel(), F32, I16, shift, fs, 0, y.data_ptr(), y.numel())
    mixer.synchronize()
    assert sn == slicing.seed_const(shift, fs, n)
    rng = np.random.default_rng(2)
    for b, e in _windows(rng, n, 1, 24):
        want, _ = oracle.mix(x[b * 8:e * 8].cpu().numpy(), F32, I16, shift, fs, samplenum=slicing.seed_const(shift, fs, b))
        assert _same(y[b * 4:e * 4].cpu().numpy(), want, I16), (b, e)
    # chunked == whole
    h = n // 2 + 1
    y2 = torch.empty((n - h) * 4, dtype=torch.uint8, device="cuda")
    mixer.mix_dev(x.data_ptr() + h * 8, (n - h) * 8, F32, I16, shift, fs, slicing.seed_const(shift, fs, h), y2.data_ptr(), y2.numel())
    mixer.synchronize()
    assert torch.equal(y[h * 4:], y2)


def test_cfg3_track_replay_i16_full_size(oracle, mixer):
    fs, secs = 1_024_000, 600
    n = secs * fs - 555                                  # 614.4 M samples, short last block
    table = _overpass_table(oracle, secs, 437_505_000, 300.0)
    shifts = dsp.replay_schedule(table, 5000, fs, I16, n * 4)
    bs = slicing.block_samples(I16)
    assert shifts.size == (n + bs - 1) // bs
    x = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
    y = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
    _fill(x, I16)
    _, _, _, stats = dsp.plan_tiles_trace(I16, I16, 0, shifts[:8 * fs // bs], bs, fs, 8 * fs)
    assert stats["column_segments"] >= 2, stats        # long-period pieces do take the COLUMN path
    sn = mixer.mix_blocks_dev(x.data_ptr(), x.numel(), I16, I16, shifts, fs, 0, y.data_ptr(), y.numel())
    mixer.synchronize()
    assert sn == slicing.seed_blocks(shifts, I16, fs, n)
    rng = np.random.default_rng(3)
    for b, e in _windows(rng, n, bs, 24):
        want, _ = oracle.mix_blocks(x[b * 4:e * 4].cpu().numpy(), I16, I16, shifts[b // bs:], fs,
                                    samplenum=slicing.seed_blocks(shifts, I16, fs, b))
        assert _same(y[b * 4:e * 4].cpu().numpy(), want, I16), (b, e)
    h = (n // 2) // bs * bs
    y2 = torch.empty((n - h) * 4, dtype=torch.uint8, device="cuda")
    mixer.mix_blocks_dev(x.data_ptr() + h * 4, (n - h) * 4, I16, I16, shifts[h // bs:], fs, slicing.seed_blocks(shifts, I16, fs, h),
                         y2.data_ptr(), y2.numel())
    mixer.synchronize()
    assert torch.equal(y[h * 4:], y2)


def test_cfg4_track_f32_one_slice_of_the_8_gpu_job(oracle, mixer):
    fs, secs, world, rank = 200_000_000, 60, 8, 3
    total = secs * fs                                    # 12 G samples in the whole job
    begin, end = slicing.slice_bounds(total, world, rank, F32)
    n = end - begin                                      # 1.5 G samples: two launches
    assert begin > 2**32 and n > 2**30
    table = _overpass_table(oracle, secs, 4_200_000_000, 30.0)
    shifts = dsp.replay_schedule(table, 0, fs, F32, total * 8)
    bs = slicing.block_samples(F32)
    seed = slicing.seed_blocks(shifts, F32, fs, begin)
    x = torch.empty(n * 8, dtype=torch.uint8, device="cuda")
    y = torch.empty(n * 8, dtype=torch.uint8, device="cuda")
    _fill(x, F32)
    sl = shifts[begin // bs:]
    sn = mixer.mix_blocks_dev(x.data_ptr(), x.numel(), F32, F32, sl, fs, seed, y.data_ptr(), y.numel())
    mixer.synchronize()
    assert sn == slicing.seed_blocks(shifts, F32, fs, end)
    rng = np.random.default_rng(4)
    for b, e in _windows(rng, n, bs, 12):
        want, _ = oracle.mix_blocks(x[b * 8:e * 8].cpu().numpy(), F32, F32, sl[b // bs:], fs,
                                    samplenum=slicing.seed_blocks(shifts, F32, fs, begin + b))
        assert _same(y[b * 8:e * 8].cpu().numpy(), want, F32), (b, e)
    # 2^24 samples straddling every slice boundary = the last 2^23 samples of a slice and the first 2^23 of the
    # next; here both ends of this slice (bench.py does the same for every slice of the job, plus the first and last second)
    import os
    from tools import workloads as W
    half = 1 << 23
    ok, checked, _ = W.check_windows(oracle, x, y, F32, F32, shifts, fs, begin, [(0, half), ((n - half) // bs * bs, n)], len(os.sched_getaffinity(0)))
    assert ok and checked >= 2 * half
    h = (n // 2) // bs * bs
    y2 = torch.empty((n - h) * 8, dtype=torch.uint8, device="cuda")
    mixer.mix_blocks_dev(x.data_ptr() + h * 8, (n - h) * 8, F32, F32, sl[h // bs:], fs, slicing.seed_blocks(shifts, F32, fs, begin + h),
                         y2.data_ptr(), y2.numel())
    mixer.synchronize()
    assert torch.equal(y[h * 8:], y2)
