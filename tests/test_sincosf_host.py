"""Host twin of the device sincosf (doppler_b200/csrc/sincosf_glibc.h compiled for the CPU by
tests/native) against this box's libm sincosf -- the function the reference reaches through
complex.c:35.  The full 2^32 sweep takes ~15 s on 8 cores (set DOPPLER_FULL_SWEEP=1); the default
run covers every exponent with a stride and every theta the BASELINE configs form."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostcheck():
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "native", "libhostcheck.so"))
    lib.hostcheck_sweep.restype = ctypes.c_uint64
    lib.hostcheck_sweep.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
    lib.hostcheck_fast_sweep.restype = ctypes.c_uint64
    lib.hostcheck_fast_sweep.argtypes = lib.hostcheck_sweep.argtypes
    lib.hostcheck_sincosf.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    return lib


def test_strided_sweep_all_exponents(hostcheck):
    full = os.environ.get("DOPPLER_FULL_SWEEP") == "1"
    stride = 1 if full else 61
    count = (2**32 + stride - 1) // stride
    bad = ctypes.c_uint32(0)
    n = hostcheck.hostcheck_sweep(0, count, stride, os.cpu_count() or 1, ctypes.byref(bad))
    assert n == 0, f"{n} mismatches vs libm, e.g. bits 0x{bad.value:08x}"


def test_range_specialised_twins_all_exponents(hostcheck):
    """db_sincosf_large / _medium / _small (the direct-evaluation fast rows of the mixer kernel): every
    pattern evaluated by the routine of its own glibc range, large-range window from its own exponent."""
    full = os.environ.get("DOPPLER_FULL_SWEEP") == "1"
    stride = 1 if full else 61
    count = (2**32 + stride - 1) // stride
    bad = ctypes.c_uint32(0)
    n = hostcheck.hostcheck_fast_sweep(0, count, stride, os.cpu_count() or 1, ctypes.byref(bad))
    assert n == 0, f"{n} mismatches vs libm, e.g. bits 0x{bad.value:08x}"
    for centre in (0x39800000, 0x3F400000, 0x42F00000, 0x7F7FFFFF):   # range boundaries, both signs
        for sign in (0, 0x80000000):
            n = hostcheck.hostcheck_fast_sweep((centre - 200_000) | sign, 400_000, 1, 2, ctypes.byref(bad))
            assert n == 0, hex(bad.value)


def test_dense_around_branch_points(hostcheck):
    # pi/4 (0x3f490fdb), 2^-12 (0x39800000), 120.0 (0x42f00000), FLT_MAX/Inf boundary, both signs
    for centre in (0x3F490FDB, 0x39800000, 0x42F00000, 0x7F7FFFFF, 0x00800000, 0x00000000):
        for sign in (0, 0x80000000):
            first = max(0, centre - 200_000) | sign
            bad = ctypes.c_uint32(0)
            n = hostcheck.hostcheck_sweep(first, 400_000, 1, 2, ctypes.byref(bad))
            assert n == 0, hex(bad.value)


def test_thetas_of_baseline_configs(hostcheck, oracle):
    """Every theta = -2*pi*(r*n) that cfg1..cfg5 can form (n over one full period)."""
    for shift, fs, period in [(-15000.0, 256000, 256), (100000.0, 10_000_000, 100), (815000.0, 2_400_000, 480),
                              (-9876.54, 1_024_000, 111_145), (7321.7, 1_024_000, 55_244), (4_000_000.5, 200_000_000, 200_000)]:
        n = np.arange(0, period + 2, dtype=np.uint32)
        r = np.float32(shift) / np.float32(fs)
        theta = (np.float32(-2.0) * np.float32(np.pi)) * (r * n.astype(np.float32))
        s_ref, c_ref = oracle.sincosf_batch(theta)
        s = ctypes.c_float()
        c = ctypes.c_float()
        step = max(1, period // 5000)
        for i in range(0, theta.size, step):
            hostcheck.hostcheck_sincosf(float(theta[i]), ctypes.byref(s), ctypes.byref(c))
            assert np.float32(s.value).tobytes() == s_ref[i].tobytes() and np.float32(c.value).tobytes() == c_ref[i].tobytes()
