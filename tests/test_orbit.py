"""Host orbit propagation for track mode (doppler_b200/csrc/orbit.cpp: TLE parser, SGP4, observer
range rate) -- the stand-in for crate gpredict / libgpredict (main.rs:141-163).  Parity with
libgpredict is UNPINNED (absent offline); the propagator is checked against the
published verification case of Spacetrack Report No. 3 and against physical invariants."""
import ctypes

import numpy as np
import pytest

from doppler_b200 import _lib

# Spacetrack Report No. 3, section 13, SGP4 test case (values printed there in single precision)
L1 = "1 88888U          80275.98708465  .00073094  13844-3  66816-4 0    87"
L2 = "2 88888  72.8435 115.9689 0086731  52.6988 110.5714 16.05824518  1058"
STR3 = {
    0.0: (2328.97048951, -5995.22076416, 1719.97067261, 2.91207230, -0.98341546, -7.09081703),
    360.0: (2456.10705566, -6071.93853760, 1222.89727783, 2.67938992, -0.44829041, -7.22879231),
    720.0: (2567.56195068, -6112.50384522, 713.96397400, 2.44024599, 0.09810869, -7.31995916),
    1080.0: (2663.09078980, -6115.48229980, 196.39640427, 2.19611958, 0.65241995, -7.36282432),
    1440.0: (2742.55133057, -6079.67144775, -326.38095856, 1.94850229, 1.21106251, -7.35619372),
}


def make(l1=L1, l2=L2, lat=58.26541, lon=26.46667, alt=76.0):
    lib = _lib.load()
    tr = ctypes.c_void_p()
    rc = lib.doppler_b200_tracker_create_from_lines(b"TEST", l1.encode(), l2.encode(), lat, lon, alt, ctypes.byref(tr))
    return lib, tr, rc


def test_sgp4_reproduces_spacetrack_report_3():
    lib, tr, rc = make()
    assert rc == 0
    for t, exp in STR3.items():
        p, v = np.zeros(3), np.zeros(3)
        assert lib.doppler_b200_tracker_teme(tr, t, p.ctypes.data, v.ctypes.data) == 0
        assert np.abs(p - np.array(exp[:3])).max() < 2e-2   # km; the report is single precision
        assert np.abs(v - np.array(exp[3:])).max() < 2e-5   # km/s
    lib.doppler_b200_tracker_destroy(tr)


def test_bad_checksum_and_deep_space_are_rejected():
    lib, tr, rc = make(l1=L1[:-1] + "3")
    assert rc != 0 and b"checksum" in lib.doppler_b200_tracker_last_error()
    # a geostationary-like mean motion (1.0027 rev/day) is a deep-space object: SDP4 not implemented
    l2 = "2 88888  72.8435 115.9689 0086731  52.6988 110.5714  1.00270000  105"
    s = sum((int(c) if c.isdigit() else (1 if c == "-" else 0)) for c in l2[:68]) % 10
    lib, tr, rc = make(l2=l2 + str(s))
    assert rc != 0 and b"deep-space" in lib.doppler_b200_tracker_last_error()


def test_range_rate_is_the_derivative_of_range_and_doppler_table_follows_main_rs():
    lib, tr, rc = make()
    assert rc == 0
    # epoch 1980 day 275.98708465 -> unix seconds
    epoch = (np.datetime64("1980-01-01") - np.datetime64("1970-01-01")) / np.timedelta64(1, "s") + (275.98708465 - 1.0) * 86400.0
    az, el, rng, rr = (ctypes.c_double() for _ in range(4))

    def obs(t):
        assert lib.doppler_b200_tracker_observe(tr, t, ctypes.byref(az), ctypes.byref(el), ctypes.byref(rng), ctypes.byref(rr)) == 0
        return az.value, el.value, rng.value, rr.value

    for t in epoch + np.array([0.0, 600.0, 4000.0, 86400.0]):
        _, e0, r0, rr0 = obs(t)
        _, _, r1, _ = obs(t + 0.5)
        _, _, rm, _ = obs(t - 0.5)
        assert abs((r1 - rm) - rr0) < 3e-4        # central difference over 1 s, km/s (JD in double: ~40 us time grain)
        assert abs(rr0) < 8.5 and -90.0 <= e0 <= 90.0
    tab = np.zeros(7)
    assert lib.doppler_b200_tracker_doppler_table(tr, float(epoch), 437_505_000, 7, tab.ctypes.data) == 7
    for s in range(7):
        _, _, _, rrs = obs(epoch + s)
        assert tab[s] == lib.doppler_b200_doppler_hz(rrs, 437_505_000)
        assert tab[s] == (rrs * 1000.0 / 299792458.0) * 437_505_000.0 * (-1.0)   # main.rs:163, same association
    lib.doppler_b200_tracker_destroy(tr)


def test_tle_from_file(tmp_path):
    f = tmp_path / "cubesat.txt"
    f.write_text("OTHER SAT\n" + L1 + "\n" + L2 + "\nSYNTHETIC TEST SAT   \r\n" + L1 + "\r\n" + L2 + "\r\n")
    lib = _lib.load()
    tr = ctypes.c_void_p()
    assert lib.doppler_b200_tracker_create(str(f).encode(), b"SYNTHETIC TEST SAT", 58.0, 26.0, 0.0, ctypes.byref(tr)) == 0
    lib.doppler_b200_tracker_destroy(tr)
    assert lib.doppler_b200_tracker_create(str(f).encode(), b"ESTCUBE 1", 58.0, 26.0, 0.0, ctypes.byref(tr)) != 0
    assert b"not found" in lib.doppler_b200_tracker_last_error()
