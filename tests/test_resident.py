"""The resident kernel of the per-block host path (mixer_kernels.cuh: mix_resident_kernel; include/doppler_b200.h:
DOPPLER_B200_TUNE_RESIDENT_IDLE_US).  The reference calls its mixer once per 8192-byte block (main.rs:49,70); here such calls are
served by one CTA that stays on the chip and takes its requests from a mailbox in pinned host memory.  Results must not depend on
which path served a block, the kernel must leave when idle and come back on demand, and nothing may hang when the context ends."""
import time

import numpy as np
import pytest

import doppler_b200
from doppler_b200 import F32, I16
from tests.oracle_lib import BUFFER_SIZE, same_bits_f32

BPS = {I16: 4, F32: 8}
pytestmark = pytest.mark.gpu


def make_input(rng, n, typ):
    if typ == I16:
        return rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype("<i2").view(np.uint8)
    return rng.uniform(-0.7, 0.7, 2 * n).astype("<f4").view(np.uint8)


def same(got, want, outtype):
    return np.array_equal(got, want) if outtype == I16 else same_bits_f32(got, want)


@pytest.fixture
def fresh():
    m = doppler_b200.Mixer(0)
    yield m
    m.close()


@pytest.mark.parametrize("intype,outtype", [(I16, I16), (I16, F32), (F32, I16), (F32, F32)])
def test_block_stream_is_served_without_a_launch_per_block(oracle, fresh, intype, outtype):
    """300 pump blocks with a new shift every few blocks (track mode's call pattern), samplenum carried: bytes equal to the
    oracle's chain, and the library launched a handful of kernels (the resident one, a table or two), not one per block."""
    rng = np.random.default_rng(11 + 2 * intype + outtype)
    fs, bs = 1_024_000, BUFFER_SIZE // BPS[intype]
    shifts = np.repeat(rng.uniform(-12000, 12000, 60).astype(np.float32), 5)
    sn_g = sn_o = 0
    before = fresh.launch_count
    for b, shift in enumerate(shifts):
        n = bs if b % 7 else int(rng.integers(1, bs))        # ragged blocks in between
        buf = make_input(rng, n, intype)
        got, sn_g = fresh.mix(buf, intype, outtype, float(shift), fs, samplenum=sn_g)
        want, sn_o = oracle.mix(buf, intype, outtype, float(shift), fs, samplenum=sn_o)
        assert sn_g == sn_o and same(got, want, outtype), b
    assert fresh.launch_count - before <= 8


def test_const_mode_blocks_with_a_phasor_table(oracle, fresh):
    """Short period (P = 256): the blocks read a phasor table built by an ordinary launch on another stream."""
    rng = np.random.default_rng(5)
    sn_g = sn_o = 0
    for b in range(50):
        buf = make_input(rng, 2048, I16)
        got, sn_g = fresh.mix(buf, I16, I16, -15000.0, 256000, samplenum=sn_g)
        want, sn_o = oracle.mix(buf, I16, I16, -15000.0, 256000, samplenum=sn_o)
        assert sn_g == sn_o and np.array_equal(got, want), b


def test_kernel_leaves_when_idle_and_returns_on_demand(oracle, fresh):
    fresh.tune(resident_idle_us=2000)
    rng = np.random.default_rng(6)
    buf = make_input(rng, 2048, I16)
    want, _ = oracle.mix(buf, I16, F32, 7321.7, 1_024_000)
    starts = []
    for pause in (0.0, 0.0, 0.05, 0.0, 0.05, 0.004, 0.0):    # longer and shorter than the time-out
        time.sleep(pause)
        before = fresh.launch_count
        got, _ = fresh.mix(buf, I16, F32, 7321.7, 1_024_000)
        assert same_bits_f32(got, want)
        starts.append(fresh.launch_count - before)
    assert starts[0] >= 1 and starts[1] == 0                 # started by the first block, reused by the second
    assert starts[2] >= 1 and starts[4] >= 1                 # gone after 50 ms of silence: a new one
    assert sum(starts) <= 6


def test_paths_agree_and_interleave(oracle, fresh):
    """Resident kernel, one zero-copy launch per block, staged pipeline and a large call in between: same bytes."""
    rng = np.random.default_rng(8)
    small = make_input(rng, 1024, F32)
    big = make_input(rng, 3_000_001, F32)
    want_s, _ = oracle.mix(small, F32, I16, 100000.0, 10_000_000, samplenum=17)
    want_b, _ = oracle.mix(big, F32, I16, 100000.0, 10_000_000)
    for idle in (20000, 0, 20000):
        fresh.tune(resident_idle_us=idle)
        got, _ = fresh.mix(small, F32, I16, 100000.0, 10_000_000, samplenum=17)
        assert np.array_equal(got, want_s)
        got, _ = fresh.mix(big, F32, I16, 100000.0, 10_000_000)
        assert np.array_equal(got, want_b)
        got, _ = fresh.mix(small, F32, I16, 100000.0, 10_000_000, samplenum=17)
        assert np.array_equal(got, want_s)
    fresh.tune(tiny_host_bytes=0)
    got, _ = fresh.mix(small, F32, I16, 100000.0, 10_000_000, samplenum=17)
    assert np.array_equal(got, want_s)


def test_context_ends_while_the_kernel_is_resident(oracle):
    """destroy() right after a block: the kernel is told to leave, nothing waits for the idle time-out."""
    rng = np.random.default_rng(9)
    buf = make_input(rng, 2048, I16)
    want, _ = oracle.mix(buf, I16, I16, 5000.0, 1_024_000)
    for _ in range(3):
        m = doppler_b200.Mixer(0)
        m.tune(resident_idle_us=5_000_000)
        got, _ = m.mix(buf, I16, I16, 5000.0, 1_024_000)
        assert np.array_equal(got, want)
        t0 = time.perf_counter()
        m.close()
        assert time.perf_counter() - t0 < 1.0


@pytest.mark.parametrize("intype,outtype", [(I16, I16), (I16, F32), (F32, I16), (F32, F32)])
def test_every_block_size_through_the_same_units(oracle, fresh, intype, outtype):
    """The result comes back as flagged 8-byte units in ONE buffer reused by every block (collect.cpp): shrinking, growing and
    ragged block sizes (group tails of 1..3 samples, the 32 KiB limit of the path and the first size beyond it) must never
    pick up a unit of an earlier block."""
    rng = np.random.default_rng(21 + 2 * intype + outtype)
    top = (32 << 10) // BPS[intype]                              # largest block the resident kernel serves
    sizes = list(range(1, 41)) + [2047, 2048, 2049, top - 1, top, top + 1, 3, top, 1, 2048, 5, top - 3, 2]
    sn_g = sn_o = 123
    before = fresh.launch_count
    for n in sizes:
        buf = make_input(rng, n, intype)
        got, sn_g = fresh.mix(buf, intype, outtype, -4321.5, 1_024_000, samplenum=sn_g)
        want, sn_o = oracle.mix(buf, intype, outtype, -4321.5, 1_024_000, samplenum=sn_o)
        assert sn_g == sn_o and same(got, want, outtype), n
    assert fresh.launch_count - before <= 8                      # (two blocks above the limit, the resident kernel, a table or two)


def test_identical_blocks_back_to_back(oracle, fresh):
    """The same bytes in and (with a table, P = 256 dividing the block) the same bytes out, block after block: only the request
    number in the units tells a new result from the previous one."""
    rng = np.random.default_rng(31)
    buf = make_input(rng, 2048, I16)
    sn_g = sn_o = 0
    for b in range(40):
        got, sn_g = fresh.mix(buf, I16, I16, -15000.0, 256000, samplenum=sn_g)
        want, sn_o = oracle.mix(buf, I16, I16, -15000.0, 256000, samplenum=sn_o)
        assert sn_g == sn_o and np.array_equal(got, want), b


def test_plans_beyond_the_request_lines_keep_the_launch(oracle, fresh):
    """A request carries at most three pieces; a small call with more (one shift per 256 bytes of input) is an ordinary launch,
    and the resident kernel serves the blocks around it."""
    rng = np.random.default_rng(41)
    fs = 1_024_000
    sn_g = sn_o = 0
    for b in range(12):
        buf = make_input(rng, 2048, I16)
        if b % 3 == 1:
            shifts = rng.uniform(-9000, 9000, 32).astype(np.float32)
            got, sn_g = fresh.mix_blocks(buf, I16, F32, shifts, fs, samplenum=sn_g, block_bytes=256)
            parts = []
            for i, s in enumerate(shifts):                        # the same thing as 32 chained calls of the reference's mixer
                part, sn_o = oracle.mix(buf[256 * i:256 * (i + 1)], I16, F32, float(s), fs, samplenum=sn_o)
                parts.append(part)
            want = np.concatenate(parts)
        else:
            got, sn_g = fresh.mix(buf, I16, F32, 2500.0 * b, fs, samplenum=sn_g)
            want, sn_o = oracle.mix(buf, I16, F32, 2500.0 * b, fs, samplenum=sn_o)
        assert sn_g == sn_o and same_bits_f32(got, want), b


def test_time_out_races(oracle, fresh):
    """An idle time-out of the order of the gap between calls: the kernel leaves and returns hundreds of times, with requests
    arriving before, during and after its last look (served by the leaving kernel, by a successor queued behind it, or by
    one started when the host finds nobody there).  Every block exactly once, in order: bytes and samplenum chain as the oracle's."""
    rng = np.random.default_rng(51)
    fresh.tune(resident_idle_us=120)
    fs = 1_024_000
    sn_g = sn_o = 0
    before = fresh.launch_count
    for b in range(500):
        n = 2048 if b % 5 else int(rng.integers(1, 2049))
        buf = make_input(rng, n, I16)
        shift = 5000.0 if b % 50 < 40 else float(rng.uniform(-9000, 9000))   # mostly the steady-state rule, some new ratios
        t_end = time.perf_counter() + float(rng.uniform(0, 250e-6))
        while time.perf_counter() < t_end:
            pass
        got, sn_g = fresh.mix(buf, I16, I16, shift, fs, samplenum=sn_g)
        want, sn_o = oracle.mix(buf, I16, I16, shift, fs, samplenum=sn_o)
        assert sn_g == sn_o and np.array_equal(got, want), b
    assert fresh.launch_count - before >= 20    # (it did leave and return)


def test_kernel_leaves_before_a_persistent_launch(oracle, fresh):
    """The persistent kernels take one CTA per SM with the whole register file; the resident CTA would make one of them wait
    for a neighbour (twice the launch time with the statically dealt lean kernel: tools/gpu/resident_coexist.py).  So a
    launch above the small-kernel limit sends it away first, and the next block brings it back."""
    rng = np.random.default_rng(61)
    fresh.tune(resident_idle_us=5_000_000)
    blk = make_input(rng, 2048, I16)
    big = make_input(rng, 5_000_001, I16)
    want_blk, _ = oracle.mix(blk, I16, I16, 5000.0, 1_024_000, samplenum=7)
    want_big, _ = oracle.mix(big, I16, I16, -15000.0, 256000)
    for _ in range(3):   # (the resident kernel starts; the second call knows the period and builds the table, which restarts it)
        fresh.mix(blk, I16, I16, 5000.0, 1_024_000, samplenum=7)
    before = fresh.launch_count
    got, _ = fresh.mix(blk, I16, I16, 5000.0, 1_024_000, samplenum=7)
    assert np.array_equal(got, want_blk) and fresh.launch_count == before          # served by the kernel on the chip
    got, _ = fresh.mix(big, I16, I16, -15000.0, 256000)
    assert np.array_equal(got, want_big)
    mid = fresh.launch_count
    got, _ = fresh.mix(blk, I16, I16, 5000.0, 1_024_000, samplenum=7)
    assert np.array_equal(got, want_blk) and fresh.launch_count == mid + 1         # a new resident kernel


def test_block_stream_shorter_than_two_periods_gets_its_table(oracle, fresh):
    """P = 1024 and blocks of 2047 / 1500 / 2048 samples: one block alone would not pay for a phasor table (fewer than two
    periods), the stream does -- built on the second block of the ratio, read by every block after it.  Bytes as the oracle's
    either way; the table shows as exactly one more launch."""
    rng = np.random.default_rng(71)
    fs = 1_024_000
    for n in (2047, 1500, 2048):
        m = doppler_b200.Mixer(0)
        m.tune(resident_idle_us=5_000_000)               # (a pause on a busy box must not show up as a kernel start below)
        try:
            sn_g = sn_o = 0
            counts = []
            for b in range(12):
                buf = make_input(rng, n, I16)
                before = m.launch_count
                got, sn_g = m.mix(buf, I16, I16, 5000.0, fs, samplenum=sn_g)
                want, sn_o = oracle.mix(buf, I16, I16, 5000.0, fs, samplenum=sn_o)
                assert sn_g == sn_o and np.array_equal(got, want), (n, b)
                counts.append(m.launch_count - before)
            assert sum(counts[3:]) == 0, counts          # steady: no launches at all
            assert 2 <= sum(counts[:3]) <= 4, counts     # the resident kernel, the table (and a restart once the arena exists)
        finally:
            m.close()
