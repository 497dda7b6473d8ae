"""Time-sliced multi-GPU behind the C ABI (include/doppler_b200.h: doppler_b200_slice_*, doppler_b200_multi_*,
doppler_b200_mix*_multi*).

CPU: the slice rule and the analytically carried samplenum against the oracle's sequential recurrence.
GPU: a device group mixes one stream as contiguous slices and must reproduce the single-stream oracle bytes and
final samplenum.  On a one-GPU box the group is [0, 0] (two contexts, two host threads, two slices on the same
device) so the slicing / seeding / threading is exercised by the driver's 1-GPU run; with more GPUs every device
gets a slice."""
import os
import subprocess

import numpy as np
import pytest

import doppler_b200
from doppler_b200 import F32, I16, dsp, slicing
from tests.oracle_lib import BUFFER_SIZE, same_bits_f32

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "doppler_b200", "bin", "doppler")
BPS = {I16: 4, F32: 8}


# ---- CPU ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("total", [0, 1, 2047, 2048, 5 * 2048 + 17, 16 * 2048])
@pytest.mark.parametrize("nslices", [1, 2, 3, 8])
def test_slice_seeds_const_match_the_oracle_recurrence(oracle, total, nslices):
    for shift, fs in ((-15000.0, 256000), (7321.7, 1_024_000), (1.0, 2_000_000_000)):
        begins, seeds = dsp.slice_seeds(5, shift, 2048, fs, total, nslices)
        assert begins[0] == 0 and begins[-1] == total and len(begins) == nslices + 1
        for i in range(nslices):
            assert begins[i] <= begins[i + 1]
            if i + 1 < nslices:
                assert begins[i + 1] % 2048 == 0
            assert (begins[i], begins[i + 1]) == slicing.slice_bounds(total, nslices, i, I16)
        for i in range(nslices + 1):
            assert seeds[i] == oracle.samplenum_advance(5, shift, fs, begins[i]), (shift, i)


def test_slice_seeds_schedule_match_the_oracle(oracle):
    fs, bs = 1_024_000, 1024                      # f32 input: 1024 samples per 8192-byte block
    rng = np.random.default_rng(5)
    shifts = np.repeat(rng.uniform(-12000, 12000, 9).astype(np.float32), 3)   # 27 blocks, 9 distinct shifts
    total = 26 * bs + 100
    x = rng.uniform(-0.5, 0.5, 2 * total).astype(np.float32).view(np.uint8)
    begins, seeds = dsp.slice_seeds(0, shifts, bs, fs, total, 4)
    for i in range(4):
        _, sn = oracle.mix_blocks(x[:begins[i] * 8], F32, F32, shifts, fs)
        assert (seeds[i] == sn) if begins[i] else (seeds[i] == 0)
        assert seeds[i] == slicing.seed_blocks(shifts, F32, fs, begins[i])
    _, sn = oracle.mix_blocks(x, F32, F32, shifts, fs)
    assert seeds[4] == sn


def test_slice_arguments_are_checked():
    with pytest.raises(doppler_b200.DopplerError):
        dsp.slice_bounds(100, 0, 0, 2048)
    with pytest.raises(doppler_b200.DopplerError):
        dsp.slice_bounds(100, 2, 2, 2048)
    with pytest.raises(doppler_b200.DopplerError):          # schedule shorter than the stream
        dsp.slice_seeds(0, np.zeros(2, dtype=np.float32), 1024, 48000, 5000, 2)


def test_multi_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(doppler_b200.DopplerError) as ei:
        doppler_b200.MultiMixer([0, 1])
    assert ei.value.code == dsp.ENODEV


# ---- GPU ------------------------------------------------------------------------------------------

def _groups():
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    groups = [[0, 0], [0, 0, 0]]
    if n >= 2:
        groups.append(list(range(n)))
    return groups


@pytest.fixture(scope="module", params=[0, 1, 2], ids=["dev0x2", "dev0x3", "all"])
def group(request):
    gs = _groups()
    if request.param >= len(gs):
        pytest.skip("one GPU on this box")
    m = doppler_b200.MultiMixer(gs[request.param])
    yield m
    m.close()


def _input(rng, n, typ):
    if typ == I16:
        return rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype("<i2").view(np.uint8)
    return rng.uniform(-0.7, 0.7, 2 * n).astype("<f4").view(np.uint8)


def _same(got, want, outtype):
    return np.array_equal(got, want) if outtype == I16 else same_bits_f32(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("intype,outtype", [(I16, I16), (I16, F32), (F32, I16), (F32, F32)])
@pytest.mark.parametrize("shift,fs", [(-15000.0, 256000), (7321.7, 1_024_000), (100000.0, 10_000_000)])
def test_mix_multi_host_matches_oracle(oracle, group, intype, outtype, shift, fs):
    rng = np.random.default_rng(11)
    for n in (0, 5, 2048, 40_000 + 3, 300_001):
        x = _input(rng, n, intype)
        got, sn = group.mix(x, intype, outtype, shift, fs, samplenum=3)
        want, sn_ref = oracle.mix(x, intype, outtype, shift, fs, samplenum=3)
        assert sn == sn_ref
        assert _same(got, want, outtype), n


@pytest.mark.gpu
@pytest.mark.parametrize("intype,outtype", [(I16, I16), (F32, F32)])
def test_mix_blocks_multi_host_matches_oracle(oracle, group, intype, outtype):
    rng = np.random.default_rng(12)
    fs = 1_024_000
    bs = BUFFER_SIZE // BPS[intype]
    nblocks = 37
    n = (nblocks - 1) * bs + bs // 3
    shifts = np.repeat(rng.uniform(-12000, 12000, 10).astype(np.float32), 4)[:nblocks]
    x = _input(rng, n, intype)
    got, sn = group.mix_blocks(x, intype, outtype, shifts, fs)
    want, sn_ref = oracle.mix_blocks(x, intype, outtype, shifts, fs)
    assert sn == sn_ref
    assert _same(got, want, outtype)
    with pytest.raises(doppler_b200.DopplerError):   # schedule too short
        group.mix_blocks(x, intype, outtype, shifts[:5], fs)


@pytest.mark.gpu
def test_mix_multi_dev_slices_on_their_devices(oracle):
    """Device-resident slices, one per device of the box (all on device 0 when there is one GPU)."""
    import torch
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0]
    m = doppler_b200.MultiMixer(devices)
    try:
        rng = np.random.default_rng(13)
        fs, shift = 10_000_000, 100000.0
        bs = BUFFER_SIZE // 8
        total = (len(devices) * 50 + 1) * bs + 77
        x = _input(rng, total, F32)
        begins, seeds = dsp.slice_seeds(0, shift, bs, fs, total, len(devices))
        xs, ys = [], []
        for d, dev in enumerate(devices):
            b, e = begins[d], begins[d + 1]
            xs.append(torch.from_numpy(x[b * 8:e * 8].copy()).to(f"cuda:{dev}"))
            ys.append(torch.empty((e - b) * 4, dtype=torch.uint8, device=f"cuda:{dev}"))
        for dev in set(devices):
            torch.cuda.synchronize(dev)
        sn = m.mix_dev([t.data_ptr() for t in xs], [t.numel() for t in xs], F32, I16, shift, fs, 0,
                       [t.data_ptr() for t in ys], [t.numel() for t in ys])
        m.synchronize()
        want, sn_ref = oracle.mix(x, F32, I16, shift, fs)
        assert sn == sn_ref == seeds[-1]
        got = np.concatenate([t.cpu().numpy() for t in ys])
        assert np.array_equal(got, want)
        assert m.launch_count >= len(devices)
        # a per-block schedule over the same slices
        nblocks = (total + bs - 1) // bs
        shifts = np.repeat(rng.uniform(-90000, 90000, 7).astype(np.float32), nblocks // 7 + 1)[:nblocks]
        ys2 = [torch.empty(t.numel() * 2, dtype=torch.uint8, device=t.device) for t in ys]
        sn = m.mix_blocks_dev([t.data_ptr() for t in xs], [t.numel() for t in xs], F32, F32, shifts, fs, 0,
                              [t.data_ptr() for t in ys2], [t.numel() for t in ys2])
        m.synchronize()
        want, sn_ref = oracle.mix_blocks(x, F32, F32, shifts, fs)
        assert sn == sn_ref
        assert same_bits_f32(np.concatenate([t.cpu().numpy() for t in ys2]), want)
    finally:
        m.close()


@pytest.mark.gpu
def test_cli_devices_flag_time_slices_every_chunk(oracle):
    import torch
    ndev = torch.cuda.device_count()
    devs = ",".join(str(d) for d in range(ndev)) if ndev >= 2 else "0,0"
    rng = np.random.default_rng(14)
    n = 2048 * 60 + 99
    x = _input(rng, n, I16)
    want, _, panicked = oracle.const_stream(x, I16, I16, -15000, 256000)
    assert not panicked
    r = subprocess.run([CLI, "const", "-s", "256000", "-i", "i16", "--shift", "-15000", "--devices", devs], input=x.tobytes(),
                       capture_output=True, timeout=180)
    assert r.returncode == 0, r.stderr
    assert r.stdout == want.tobytes()
    assert b"time slices" in r.stderr
