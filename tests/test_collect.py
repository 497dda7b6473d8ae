"""Host side of the resident kernel's fence-free hand-over (doppler_b200/csrc/collect.cpp): the device writes its result as
8-byte units {word, request number}; the host takes a word when its neighbour shows the number.  CPU-only: the units are
written here, in the orders a device may deliver them."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "native", "libhostcheck.so"))
    for f in (L.hostcheck_collect, L.hostcheck_collect_scalar):
        f.restype = ctypes.c_size_t
        f.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
    return L


def _units(words, flags):
    u = np.empty(2 * words.size, np.uint32)
    u[0::2] = words
    u[1::2] = flags
    return u


@pytest.mark.parametrize("which", ["hostcheck_collect", "hostcheck_collect_scalar"])
@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 63, 64, 2048, 2049, 16383])
@pytest.mark.parametrize("misalign", [0, 1, 2])
def test_complete_result_is_copied_out(lib, which, n, misalign):
    rng = np.random.default_rng(n)
    words = rng.integers(0, 2**32, n, dtype=np.uint32)
    u = _units(words, np.full(n, 77, np.uint32))
    raw = np.full(4 * n + 16, 0xAB, np.uint8)   # the caller's buffer: any alignment, nothing written beyond the result
    out = raw[misalign:]
    got = getattr(lib, which)(u.ctypes.data, 77, out.ctypes.data, 0, n)
    assert got == n
    assert np.array_equal(out[:4 * n].view(np.uint8), words.view(np.uint8))
    assert np.all(raw[misalign + 4 * n:] == 0xAB) and np.all(raw[:misalign] == 0xAB)


@pytest.mark.parametrize("which", ["hostcheck_collect", "hostcheck_collect_scalar"])
def test_stops_at_the_first_missing_word_and_resumes(lib, which):
    n = 2048
    rng = np.random.default_rng(1)
    words = rng.integers(0, 2**32, n, dtype=np.uint32)
    stale = rng.integers(0, 2**32, n, dtype=np.uint32)
    fn = getattr(lib, which)
    for hole in [0, 1, 5, 8, 15, 16, 1000, 2040, 2047]:
        flags = np.full(n, 9, np.uint32)
        flags[hole] = 8                           # the previous request's unit is still there
        data = words.copy()
        data[hole] = stale[hole]
        u = _units(data, flags)
        out = np.zeros(n, np.uint32)
        got = fn(u.ctypes.data, 9, out.ctypes.data, 0, n)
        assert got <= hole and hole - got < 8      # (the vector path works in lines of eight units)
        assert np.array_equal(out[:got], words[:got])
        u[2 * hole], u[2 * hole + 1] = words[hole], 9   # it arrives
        assert fn(u.ctypes.data, 9, out.ctypes.data, got, n) == n
        assert np.array_equal(out, words)


def test_units_arriving_in_any_order(lib):
    """Units land in arbitrary order; the collector never hands out a word whose flag is not the request's."""
    n = 1000
    rng = np.random.default_rng(2)
    old = rng.integers(0, 2**32, n, dtype=np.uint32)
    new = rng.integers(0, 2**32, n, dtype=np.uint32)
    u = _units(old, np.full(n, 41, np.uint32))
    out = np.zeros(n, np.uint32)
    got = 0
    order = rng.permutation(n)
    for step, k in enumerate(order):
        u[2 * k], u[2 * k + 1] = new[k], 42
        if step % 37 == 0 or step == n - 1:
            now = lib.hostcheck_collect(u.ctypes.data, 42, out.ctypes.data, got, n)
            assert now >= got
            assert np.array_equal(out[:now], new[:now])
            got = now
    assert got == n


@pytest.mark.parametrize("first_seq", [1, 0xFFFFFF00])
def test_hand_over_against_a_software_device(lib, first_seq):
    """The product's host code (post + collect) against a thread that plays the resident kernel's side of the protocol and is
    harsher than the device: it reads the request lines word by word in a random order and writes the result units in a random
    order.  Every request is taken whole, exactly once and in order, and no stale word is ever handed out -- also across the
    wrap of the 32-bit request number."""
    lib.hostcheck_handover.restype = ctypes.c_int
    lib.hostcheck_handover.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32]
    assert lib.hostcheck_handover(3000, 2048, 12345 + first_seq % 7, first_seq) == 0


def test_request_layout(lib):
    """dcollect::post against the layout the kernel reads (mixer_kernels.cuh: RtMailbox, rt_verdict): four 32-byte sectors, seven
    payload words each in order, the request number as every sector's eighth word, zeros behind the payload."""
    lib.hostcheck_post.restype = None
    lib.hostcheck_post.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int]
    payload = np.arange(1000, 1026, dtype=np.uint32)            # head, nsamples, three pieces of eight words
    box = np.full(32, 0xDEADBEEF, np.uint32)
    lib.hostcheck_post(box.ctypes.data, 4, 77, payload.ctypes.data, payload.size)
    assert list(box[7::8]) == [77, 77, 77, 77]
    body = np.concatenate([box[8 * t:8 * t + 7] for t in range(4)])
    assert np.array_equal(body[:26], payload) and not body[26:].any()
    # lane -> payload index as the kernel computes it: (lane >> 3) * 7 + (lane & 7) for the lanes that are not tags
    for lane in range(32):
        if lane & 7 != 7:
            w = (lane >> 3) * 7 + (lane & 7)
            assert box[lane] == (payload[w] if w < 26 else 0)
