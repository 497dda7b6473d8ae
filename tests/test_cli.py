"""The `doppler` CLI (doppler_b200/csrc/doppler_cli.cpp): argv contract of the reference's
src/usage.rs on the CPU, and byte-exact stdout against the oracle's stream drivers (restating
src/main.rs:102-119 and :155-184) on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from tests.oracle_lib import BUFFER_SIZE, F32, I16

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "doppler_b200", "bin", "doppler")

L1 = "1 88888U          80275.98708465  .00073094  13844-3  66816-4 0    87"
L2 = "2 88888  72.8435 115.9689 0086731  52.6988 110.5714 16.05824518  1058"


def run(args, stdin=b"", timeout=120):
    return subprocess.run([CLI] + args, input=stdin, capture_output=True, timeout=timeout)


def tone_i16(n, fs, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    sig = 0.25 * np.exp(2j * np.pi * 15000.0 / fs * t) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty(2 * n, dtype="<i2")
    iq[0::2] = np.clip(np.round(sig.real * 32767), -32768, 32767)
    iq[1::2] = np.clip(np.round(sig.imag * 32767), -32768, 32767)
    return iq.view(np.uint8)


# ---- argv contract (no GPU needed: every one of these exits before the device is opened) -------

def test_no_subcommand_exits_1():
    r = run([])
    assert r.returncode == 1 and b"no arguments provided, try with doppler -h" in r.stderr   # usage.rs:330-333


def test_help_and_version():
    assert run(["--help"]).returncode == 0
    assert b"--shift" in run(["const", "--help"]).stdout
    assert b"--tlename" in run(["track", "-h"]).stdout
    assert run(["--version"]).stdout.startswith(b"doppler ")


@pytest.mark.parametrize("args", [
    ["const", "-s", "256000", "-i", "i16"],                       # --shift required (usage.rs:153-157)
    ["const", "-i", "i16", "--shift", "5"],                       # --samplerate required
    ["const", "-s", "256000", "--shift", "5"],                    # --intype required
    ["const", "-s", "256000", "-i", "u8", "--shift", "5"],        # possible_values
    ["const", "-s", "abc", "-i", "i16", "--shift", "5"],          # value_t_or_exit!
    ["const", "-s", "256000", "-i", "i16", "--shift", "5.5"],     # i32
    ["const", "-s", "256000", "-i", "i16", "--shift", "3000000000"],
    ["const", "-s", "256000", "-i", "i16", "--shift", "5", "--bogus", "1"],
    ["track", "-s", "256000", "-i", "i16", "--tlename", "X", "--location", "lat=1,lon=2,alt=3", "--frequency", "1"],   # --tlefile
    ["track", "-s", "256000", "-i", "i16", "--tlefile", "f", "--tlename", "X", "--location", "lat=1,lon=2", "--frequency", "1"],
    ["track", "-s", "256000", "-i", "i16", "--tlefile", "f", "--tlename", "X", "--location", "lat=a,lon=2,alt=3", "--frequency", "1"],
    ["track", "-s", "256000", "-i", "i16", "--tlefile", "f", "--tlename", "X", "--location", "lat=1,lon=2,alt=3", "--frequency", "1",
     "--time", "2015-01-22 09:07:16"],                             # usage.rs:302-311
    ["track", "-s", "256000", "-i", "i16", "--tlefile", "/nonexistent", "--tlename", "X", "--location", "lat=1,lon=2,alt=3",
     "--frequency", "1", "--time", "2015-01-22T09:07:16"],         # main.rs:141-147
])
def test_bad_arguments_exit_1(args):
    r = run(args)
    assert r.returncode == 1, r.stderr
    assert r.stdout == b""


def test_tle_errors_exit_1(tmp_path):
    f = tmp_path / "t.txt"
    f.write_text("SAT\n" + L1[:-1] + "0\n" + L2 + "\n")
    base = ["track", "-s", "256000", "-i", "i16", "--tlefile", str(f), "--location", "lat=1,lon=2,alt=3", "--frequency", "1",
            "--time", "2015-01-22T09:07:16"]
    r = run(base + ["--tlename", "SAT"])
    assert r.returncode == 1 and b"checksum" in r.stderr
    r = run(base + ["--tlename", "NOPE"])
    assert r.returncode == 1 and b"not found" in r.stderr


# ---- stream parity (GPU) ------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("n", [256000, 255999, 2048 * 3, 0, 5])
def test_const_cfg1_matches_reference_stream(oracle, n):
    """BASELINE configs[0]: 1 s of i16 IQ @ 256 ksps, --shift -15000 (and the short-last-block variants)."""
    x = tone_i16(n, 256000, 20150122)
    want, _, panicked = oracle.const_stream(x, I16, I16, -15000, 256000)
    assert not panicked
    r = run(["const", "-s", "256000", "-i", "i16", "--shift", "-15000"], x.tobytes())
    assert r.returncode == 0, r.stderr
    assert r.stdout == want.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("intype,outtype,flags", [(F32, I16, ["-i", "f32", "-o", "i16"]), (I16, F32, ["--intype=i16", "--outtype=f32"]),
                                                  (F32, F32, ["-if32"])])
def test_const_type_pairs_and_flag_spellings(oracle, intype, outtype, flags):
    rng = np.random.default_rng(3)
    n = 70_001
    x = (rng.uniform(-0.7, 0.7, 2 * n).astype("<f4") if intype == F32 else rng.integers(-32768, 32768, 2 * n).astype("<i2")).view(np.uint8)
    want, _, panicked = oracle.const_stream(x, intype, outtype, 100000, 10_000_000)
    assert not panicked
    r = run(["const", "--samplerate", "10000000", "--shift", "100000"] + flags, x.tobytes())
    assert r.returncode == 0, r.stderr
    got = np.frombuffer(r.stdout, dtype=np.uint8)
    if outtype == I16:
        assert np.array_equal(got, want)
    else:
        from tests.oracle_lib import same_bits_f32
        assert same_bits_f32(got, want)


@pytest.mark.gpu
def test_misaligned_tail_panics_after_writing_the_full_blocks(oracle):
    """dsp.rs:87: the converter's assert fires on the short last block; earlier blocks are already out."""
    x = tone_i16(2048 * 2 + 10, 256000, 1).tobytes() + b"\x01\x02"
    want, _, panicked = oracle.const_stream(np.frombuffer(x, dtype=np.uint8), I16, I16, -15000, 256000)
    assert panicked
    r = run(["const", "-s", "256000", "-i", "i16", "--shift", "-15000"], x)
    assert r.returncode == 101 and b"panicked" in r.stderr
    assert r.stdout == want.tobytes()


@pytest.mark.gpu
def test_track_replay_from_doppler_table_matches_reference_stream(oracle, tmp_path):
    fs, secs = 1_024_000, 4
    t = np.arange(secs + 2, dtype=np.float64)
    rr = 7.5 * 7.5 * (t - 2.0) / np.sqrt(700.0 ** 2 + (7.5 * (t - 2.0)) ** 2)
    table = np.array([oracle.doppler_hz(v, 437_505_000) for v in rr])
    f = tmp_path / "doppler.txt"
    f.write_text("\n".join(repr(float(v)) for v in table))
    x = tone_i16(secs * fs - 123, fs, 1024000)
    want, _, _, panicked = oracle.track_replay_stream(x, I16, I16, table, 5000, fs)
    assert not panicked
    r = run(["track", "-s", str(fs), "-i", "i16", "--doppler-table", str(f), "--offset", "5000", "--time", "2015-01-22T09:07:16"], x.tobytes())
    assert r.returncode == 0, r.stderr
    assert r.stdout == want.tobytes()


@pytest.mark.gpu
def test_track_replay_with_tle_matches_library_schedule(oracle, tmp_path):
    """README's replay command line (synthetic element set: Spacetrack Report 3's test satellite, NOT
    a real ESTCube-1 TLE).  CLI stdout == mix_blocks fed by the tracker's per-second Doppler table."""
    import ctypes
    from doppler_b200 import _lib, dsp
    import doppler_b200
    f = tmp_path / "cubesat.txt"
    f.write_text("SYNTHETIC TEST SAT\n" + L1 + "\n" + L2 + "\n")
    fs, secs = 256000, 7
    x = tone_i16(secs * fs + 77, fs, 9)
    lib = _lib.load()
    tr = ctypes.c_void_p()
    assert lib.doppler_b200_tracker_create(str(f).encode(), b"SYNTHETIC TEST SAT", 58.26541, 26.46667, 76.0, ctypes.byref(tr)) == 0
    start = (np.datetime64("1980-10-02T00:10:00") - np.datetime64("1970-01-01T00:00:00")) / np.timedelta64(1, "s")
    table = np.zeros(secs + 2)
    lib.doppler_b200_tracker_doppler_table(tr, float(start), 437_505_000, table.size, table.ctypes.data)
    lib.doppler_b200_tracker_destroy(tr)
    assert np.ptp(table) > 1.0   # the Doppler actually moves over the recording
    shifts = dsp.replay_schedule(table, -2500, fs, I16, x.size)
    m = doppler_b200.Mixer(0)
    want, _ = m.mix_blocks(x, I16, I16, shifts, fs)
    m.close()
    want_o, _, _, _ = oracle.track_replay_stream(x, I16, I16, table, -2500, fs)
    assert np.array_equal(want, want_o)
    r = run(["track", "-s", str(fs), "-i", "i16", "--tlefile", str(f), "--tlename", "SYNTHETIC TEST SAT", "--location",
             "lat=58.26541,lon=26.46667,alt=76", "--frequency", "437505000", "--offset", "-2500", "--time", "1980-10-02T00:10:00"], x.tobytes())
    assert r.returncode == 0, r.stderr
    assert r.stdout == want.tobytes()
    assert b"range rate" in r.stderr   # main.rs:167-175 telemetry every 5 s of stream time


def _synthetic_estcube_tle():
    """A SYNTHETIC, checksum-valid element set shaped like ESTCube-1's orbit (660 km sun-synchronous, 98.1 deg) with an
    epoch next to the README's recording time.  NOT a real ESTCube-1 TLE -- none is available offline."""
    def ck(line):
        return line + str(sum((int(c) if c.isdigit() else (1 if c == "-" else 0)) for c in line[:68]) % 10)
    l1 = ck("1 39161U 13021C   15021.50000000  .00001200  00000-0  20000-3 0  999")
    l2 = ck("2 39161  98.1300  99.5000 0010000 200.0000 160.0000 14.69000000 9000")
    assert len(l1) == 69 and len(l2) == 69
    return "ESTCUBE 1\n" + l1 + "\n" + l2 + "\n"


@pytest.mark.gpu
def test_readme_recording_command_with_a_synthetic_estcube_tle(oracle, tmp_path):
    """README.md:59 verbatim (`--tlename 'ESTCUBE 1' --location lat=58.26541,lon=26.46667,alt=76 --frequency 437505000
    --offset -2500 --time 2015-01-22T09:07:16`, 256 ksps i16) on a synthetic element set: the CLI's stdout must equal the
    oracle's replay driver fed with the library's per-second Doppler table, and the Doppler must be that of a LEO pass."""
    import ctypes
    from doppler_b200 import _lib
    f = tmp_path / "cubesat.txt"
    f.write_text("SOME OTHER SAT\n" + L1 + "\n" + L2 + "\n" + _synthetic_estcube_tle())
    fs, secs = 256000, 12
    x = tone_i16(secs * fs + 1234, fs, 22)
    lib = _lib.load()
    tr = ctypes.c_void_p()
    assert lib.doppler_b200_tracker_create(str(f).encode(), b"ESTCUBE 1", 58.26541, 26.46667, 76.0, ctypes.byref(tr)) == 0
    start = (np.datetime64("2015-01-22T09:07:16") - np.datetime64("1970-01-01T00:00:00")) / np.timedelta64(1, "s")
    table = np.zeros(secs + 2)
    lib.doppler_b200_tracker_doppler_table(tr, float(start), 437_505_000, table.size, table.ctypes.data)
    lib.doppler_b200_tracker_destroy(tr)
    assert np.abs(table).max() < 11_000.0 and np.ptp(table) > 0.05     # |v_r| < 7.5 km/s at 437.5 MHz; it moves
    want, _, shifts, panicked = oracle.track_replay_stream(x, I16, I16, table, -2500, fs)
    assert not panicked
    r = run(["track", "-s", "256000", "-i", "i16", "--tlefile", str(f), "--tlename", "ESTCUBE 1", "--location",
             "lat=58.26541,lon=26.46667,alt=76", "--frequency", "437505000", "--offset", "-2500", "--time", "2015-01-22T09:07:16"], x.tobytes())
    assert r.returncode == 0, r.stderr
    assert r.stdout == want.tobytes()
    assert b"propagator" in r.stderr and b"SGP4" in r.stderr and b"doppler@437.505 MHz" in r.stderr


@pytest.mark.gpu
def test_live_pipe_latency_and_trickled_input(oracle):
    """A live producer (rtl_fm at ~1 Msps) delivers a few blocks at a time: the pump must hand back what has
    arrived within milliseconds (the reference: one 8192-byte block), not wait for a 32 MiB chunk -- and the
    bytes must equal the one-shot result however the input is sliced into writes."""
    import select
    import time
    n = 2048 * 40 + 777                                  # 40 blocks + a short last one
    x = tone_i16(n, 1_024_000, 7).tobytes()
    want, _, _ = oracle.const_stream(np.frombuffer(x, dtype=np.uint8), I16, I16, 5000, 1_024_000)
    p = subprocess.Popen([CLI, "const", "-s", "1024000", "-i", "i16", "--shift", "5000"], stdin=subprocess.PIPE,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, bufsize=0)
    got = bytearray()
    fd = p.stdout.fileno()
    # first 3 blocks, then silence: their output must arrive while stdin is still open
    p.stdin.write(x[:3 * BUFFER_SIZE])
    deadline = time.time() + 30.0                        # generous: includes CUDA start-up
    while len(got) < 3 * BUFFER_SIZE and time.time() < deadline:
        if select.select([fd], [], [], 0.2)[0]:
            got += os.read(fd, 1 << 20)
    assert len(got) == 3 * BUFFER_SIZE, "output of the blocks already delivered did not arrive while stdin was idle"
    # now a warm pump: one block at a time, each answered promptly
    t_worst = 0.0
    for b in range(3, 10):
        t0 = time.time()
        p.stdin.write(x[b * BUFFER_SIZE:(b + 1) * BUFFER_SIZE])
        need = (b + 1) * BUFFER_SIZE
        while len(got) < need and time.time() < t0 + 5.0:
            if select.select([fd], [], [], 0.05)[0]:
                got += os.read(fd, 1 << 20)
        assert len(got) == need
        t_worst = max(t_worst, time.time() - t0)
    assert t_worst < 0.25, f"block latency {t_worst * 1e3:.1f} ms"
    # the rest in odd-sized writes that do not respect block boundaries
    k = 10 * BUFFER_SIZE
    for step in (1, 8191, 8193, 30_000, 100_000):
        p.stdin.write(x[k:k + step])
        k += step
        time.sleep(0.002)
    p.stdin.write(x[k:])
    p.stdin.close()
    while True:
        chunk = os.read(fd, 1 << 20)
        if not chunk:
            break
        got += chunk
    assert p.wait(timeout=30) == 0
    assert bytes(got) == want.tobytes()


@pytest.mark.gpu
def test_track_realtime_mode_shifts_without_changing_magnitudes(tmp_path):
    """`doppler track` without --time (main.rs:186-206): Doppler at the wall clock, block by block.  The shift is
    not reproducible, but a frequency shift never changes |z|: every output magnitude equals the input's within the
    i16 truncation, and the stream length is kept."""
    f = tmp_path / "cubesat.txt"
    f.write_text("SYNTHETIC TEST SAT\n" + L1 + "\n" + L2 + "\n")
    fs = 256000
    n = 2048 * 20 + 55
    t = np.arange(n)
    sig = 0.25 * np.exp(2j * np.pi * 15000.0 / fs * t)
    iq = np.empty(2 * n, dtype="<i2")
    iq[0::2] = np.round(sig.real * 32767)
    iq[1::2] = np.round(sig.imag * 32767)
    r = run(["track", "-s", str(fs), "-i", "i16", "--tlefile", str(f), "--tlename", "SYNTHETIC TEST SAT", "--location",
             "lat=58.26541,lon=26.46667,alt=76", "--frequency", "437505000"], iq.tobytes())
    assert r.returncode == 0, r.stderr
    out = np.frombuffer(r.stdout, dtype="<i2").astype(np.float64)
    assert out.size == iq.size
    mag_in = np.hypot(iq[0::2].astype(np.float64), iq[1::2].astype(np.float64))
    mag_out = np.hypot(out[0::2], out[1::2])
    assert np.abs(mag_out - mag_in).max() < 3.0
    assert not np.array_equal(out, iq.astype(np.float64))   # it did shift
    # (the telemetry of main.rs:191-199 only appears once a second of wall time has passed)
