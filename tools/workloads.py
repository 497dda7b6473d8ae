"""tools/workloads.py -- the BASELINE.json configs as code, shared by bench.py, tools/sweep.py,
tools/track_scaling.py and the full-size GPU tests.

cfg1  const i16 @ 256 ksps, --shift -15000, 1 s            (CLI plumbing, CPU-runnable)
cfg2  const f32->i16 @ 10 Msps, --shift 100000             (the bench headline)
cfg3  track replay i16 @ 1.024 Msps, 600 s, analytic overpass (f_tx 437.505 MHz, +5 kHz offset)
cfg4  track f32->f32 @ 200 Msps, 60 s overpass (f_tx 4.2 GHz), 12 G samples cut into 8 time slices
cfg5  const i16->i16, 1 s buffers at 256 k .. 2 G sps with r = -15/256, plus one irregular ratio

Nothing here touches the oracle except the `cpu_*` / `check_*` helpers, which are the checker / CPU column.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from doppler_b200 import F32, I16, dsp, slicing  # noqa: E402

BPS = {I16: 4, F32: 8}
NAME = {I16: "i16", F32: "f32"}

CFG3 = {"fs": 1_024_000, "secs": 600, "ftx": 437_505_000, "tc": 300.0, "offset": 5000, "intype": I16, "outtype": I16}
CFG4 = {"fs": 200_000_000, "secs": 60, "ftx": 4_200_000_000, "tc": 30.0, "offset": 0, "intype": F32, "outtype": F32, "slices": 8}
CFG5_RATES = [256_000, 1_024_000, 10_000_000, 100_000_000, 1_000_000_000, 2_000_000_000]
CFG5_IRREGULAR = {"fs": 1_024_000, "shift": 7321.0}


def cfg5_shift(fs):
    """r = -15/256 at every rate (P = 256); the largest, -117 187 500 Hz, fits the CLI's i32 --shift."""
    return -15000.0 * fs / 256000.0


def overpass_doppler_table(secs, ftx, tc):
    """doppler_hz per whole second of an analytic overpass (v = 7.5 km/s, closest approach 700 km at t = tc), formed
    exactly as the reference forms it from a range rate (main.rs:163)."""
    t = np.arange(secs + 2, dtype=np.float64)
    v, d = 7500.0, 700e3
    rr_km_s = v * v * (t - tc) / np.sqrt(d * d + (v * (t - tc)) ** 2) / 1000.0
    return np.array([dsp.doppler_hz(x, ftx) for x in rr_km_s])


def overpass_shifts(fs, secs, ftx, tc, offset, intype, nsamples):
    """Per-block f32 shift schedule of the reference's replay driver (main.rs:155-184) for that overpass."""
    return dsp.replay_schedule(overpass_doppler_table(secs, ftx, tc), offset, fs, intype, nsamples * BPS[intype])


def cfg_schedule(cfg, nsamples=None):
    n = cfg["secs"] * cfg["fs"] if nsamples is None else nsamples
    return overpass_shifts(cfg["fs"], cfg["secs"], cfg["ftx"], cfg["tc"], cfg["offset"], cfg["intype"], n)


def fill_device(x, typ):
    """Synthetic IQ on the device: uniform [-0.7, 0.7) floats / uniform +-20000 int16 (chunked: randint goes through int64)."""
    import torch
    if typ == F32:
        x.view(torch.float32).uniform_(-0.7, 0.7)
    else:
        v = x.view(torch.int16)
        step = 1 << 28
        for k in range(0, v.numel(), step):
            v[k:k + step].copy_(torch.randint(-20000, 20000, (min(step, v.numel() - k),), device=x.device, dtype=torch.int16))
    torch.cuda.synchronize(x.device)


def same_bytes(got, want, outtype):
    if outtype == I16:
        return bool(np.array_equal(got, want))
    g, w = got.view(np.uint32), want.view(np.uint32)
    if g.shape != w.shape:
        return False
    nan = np.isnan(got.view(np.float32)) & np.isnan(want.view(np.float32))
    return bool(np.all((g == w) | nan))


def check_windows(oracle, x_dev, y_dev, intype, outtype, shifts, fs, begin0, windows, threads, const_shift=None):
    """Bit-compares windows [b, e) (samples, relative to the device buffers; b on a pump-block boundary) of a mixed
    stream with the oracle run on the same input window.  The oracle is seeded with the ANALYTIC samplenum of the
    window's first sample (stream position begin0 + b) and carries it on by the sequential recurrence; its state at
    the window's end must equal the analytic one there (the chain between windows is checked, not assumed).
    Returns (ok, samples checked, oracle seconds)."""
    ib, ob = BPS[intype], BPS[outtype]
    bs = slicing.block_samples(intype)
    ok, total, secs = True, 0, 0.0
    for b, e in windows:
        xin = x_dev[b * ib:e * ib].cpu().numpy()
        got = y_dev[b * ob:e * ob].cpu().numpy()
        if const_shift is not None:
            seed = dsp.samplenum_advance(0, const_shift, fs, begin0 + b)
            seed_end = dsp.samplenum_advance(0, const_shift, fs, begin0 + e)
            sh = np.full((e - b + bs - 1) // bs, const_shift, dtype=np.float32)
        else:
            seed = slicing.seed_blocks(shifts, intype, fs, begin0 + b)
            seed_end = slicing.seed_blocks(shifts, intype, fs, begin0 + e)
            sh = shifts[(begin0 + b) // bs:]
        t, want, sn = oracle.bench_blocks(xin, e - b, intype, outtype, sh, fs, threads, samplenum=seed)
        secs += t
        total += e - b
        ok = ok and same_bytes(got, want, outtype) and sn == seed_end
    return ok, total, secs


def cpu_rate(oracle, intype, outtype, fs, nsamples, threads, shift=None, shifts=None, seed=0, rng_seed=1):
    """Msamples/s of the oracle (the reference's CPU path) on `nsamples` synthetic samples in memory."""
    rng = np.random.default_rng(rng_seed)
    if intype == F32:
        x = rng.uniform(-0.7, 0.7, 2 * nsamples).astype(np.float32).view(np.uint8)
    else:
        x = rng.integers(-20000, 20000, 2 * nsamples, dtype=np.int32).astype(np.int16).view(np.uint8)
    if shifts is None:
        t, _ = oracle.bench_const(x, nsamples, intype, outtype, shift, fs, threads)
    else:
        t, _, _ = oracle.bench_blocks(x, nsamples, intype, outtype, shifts, fs, threads, samplenum=seed)
    return nsamples / t / 1e6


def cpu_columns(oracle, threads, budget_s=1.0):
    """The CPU columns of the `configs` block: the oracle on a bounded sample of every config, all host threads and
    one thread (the reference is single-threaded).  A few seconds in total."""
    out = {}

    def both(intype, outtype, fs, shift=None, shifts=None, seed=0, cap=1 << 62):
        if shifts is not None:   # never more samples than the schedule covers (many-core boxes ask for a lot)
            cap = min(cap, (len(shifts) - 1) * (8192 // BPS[intype]))
        one = cpu_rate(oracle, intype, outtype, fs, min(1 << 20, cap) // 2048 * 2048, 1, shift, shifts, seed)
        n = int(min(cap, max(1 << 21, one * 1e6 * budget_s * threads * 0.6))) // 2048 * 2048
        allc = cpu_rate(oracle, intype, outtype, fs, n, threads, shift, shifts, seed)
        return {"cpu_msps_1core": one, "cpu_msps_allcores": allc, "cpu_cores": threads, "cpu_sample": n}

    out["cfg1"] = both(I16, I16, 256_000, shift=-15000.0, cap=1 << 22)
    s3 = cfg_schedule(CFG3)
    b3 = 200 * CFG3["fs"] // 2048                       # a stretch 200 s into the pass (long-period shifts)
    out["cfg3"] = both(I16, I16, CFG3["fs"], shifts=s3[b3:], seed=slicing.seed_blocks(s3, I16, CFG3["fs"], b3 * 2048))
    s4 = cfg_schedule(CFG4)
    b4 = 20 * CFG4["fs"] // 1024
    out["cfg4"] = both(F32, F32, CFG4["fs"], shifts=s4[b4:], seed=slicing.seed_blocks(s4, F32, CFG4["fs"], b4 * 1024))
    out["cfg5"] = {}
    for fs in CFG5_RATES:                                     # r = -15/256 at every rate; the sample is at most the 1 s buffer
        out["cfg5"][str(fs)] = both(I16, I16, fs, shift=cfg5_shift(fs), cap=fs)
    out["cfg5"]["irregular"] = both(I16, I16, CFG5_IRREGULAR["fs"], shift=CFG5_IRREGULAR["shift"])
    return out


class Timer:
    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.s = time.perf_counter() - self.t0
