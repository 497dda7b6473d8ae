"""Python mirror of the reference's ``dsp`` module (/root/reference/src/dsp.rs) over the C ABI.

Same names, same argument meaning, same failure behaviour (the reference's ``assert!`` panics
become :class:`DopplerError` with ``code == EALIGN``).  ``samplenum`` -- a ``&mut u32`` in the
reference (src/dsp.rs:117, state at src/main.rs:60) -- is passed in and returned.
"""
import ctypes

import numpy as np

from . import _lib

I16 = 0  # usage.rs:39-42 DataType::I16
F32 = 1  # usage.rs:39-42 DataType::F32
BUFFER_SIZE = 8192  # main.rs:49

OK, EINVAL, EALIGN, ECAP, ECUDA, ENODEV, ENOMEM, ELIBM = range(8)
_BPS = {I16: 4, F32: 8}


class DopplerError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"doppler_b200 error {code}: {msg}")
        self.code = code


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a.size else ctypes.c_void_p(0)


def _as_bytes_array(buf):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1)
    return np.ascontiguousarray(a)


class Mixer:
    """One context per GPU (include/doppler_b200.h: doppler_b200_create)."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._ctx = ctypes.c_void_p()
        rc = self._lib.doppler_b200_create(int(device), ctypes.byref(self._ctx))
        if rc != OK:
            raise DopplerError(rc, self._lib.doppler_b200_last_error(None).decode())
        self.device = device

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.doppler_b200_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise DopplerError(rc, self._lib.doppler_b200_last_error(self._ctx).decode())

    @property
    def launch_count(self):
        return int(self._lib.doppler_b200_launch_count(self._ctx))

    def synchronize(self):
        self._check(self._lib.doppler_b200_synchronize(self._ctx))

    def tune(self, small_max_samples=None, tiny_host_bytes=None, seg_variant=None, max_claim=None, decim_variant=None, decim_stage_slots=None, resident_idle_us=None):
        """Thresholds between code paths (doppler_b200_tune); results are identical on every path."""
        if resident_idle_us is not None:
            self._check(self._lib.doppler_b200_tune(self._ctx, 7, int(resident_idle_us)))
        if decim_stage_slots is not None:
            self._check(self._lib.doppler_b200_tune(self._ctx, 6, int(decim_stage_slots)))
        if decim_variant is not None:
            self._check(self._lib.doppler_b200_tune(self._ctx, 5, int(decim_variant)))
        if max_claim is not None:
            self._check(self._lib.doppler_b200_tune(self._ctx, 4, int(max_claim)))
        if seg_variant is not None:
            self._check(self._lib.doppler_b200_tune(self._ctx, 3, int(seg_variant)))
        if small_max_samples is not None:
            self._check(self._lib.doppler_b200_tune(self._ctx, 1, int(small_max_samples)))
        if tiny_host_bytes is not None:
            self._check(self._lib.doppler_b200_tune(self._ctx, 2, int(tiny_host_bytes)))

    # -- reference functions -------------------------------------------------------------
    def convert_iqi16_to_complex(self, inbuf):
        """dsp.rs:85-99."""
        a = _as_bytes_array(inbuf)
        out = np.empty(a.size // 4, dtype=np.complex64)
        self._check(self._lib.doppler_b200_convert_iqi16_to_complex(self._ctx, _ptr(a), a.size, _ptr(out)))
        return out

    def convert_iqf32_to_complex(self, inbuf):
        """dsp.rs:101-115."""
        a = _as_bytes_array(inbuf)
        out = np.empty(a.size // 8, dtype=np.complex64)
        self._check(self._lib.doppler_b200_convert_iqf32_to_complex(self._ctx, _ptr(a), a.size, _ptr(out)))
        return out

    def shift_frequency(self, inbuf, samplenum, shift_hz, samplerate):
        """dsp.rs:117-134.  Returns (output complex64 array, new samplenum)."""
        a = np.ascontiguousarray(inbuf, dtype=np.complex64)
        out = np.empty_like(a)
        sn = ctypes.c_uint32(samplenum)
        self._check(self._lib.doppler_b200_shift_frequency(self._ctx, _ptr(a), a.size, ctypes.byref(sn),
                                                           ctypes.c_float(shift_hz), int(samplerate), _ptr(out)))
        return out, sn.value

    # -- fused path ----------------------------------------------------------------------
    def mix(self, inbuf, intype, outtype, shift_hz, samplerate, samplenum=0):
        """convert -> shift_frequency -> egress (main.rs:65-94) in one pass.  Returns (bytes array, samplenum)."""
        a = _as_bytes_array(inbuf)
        out = np.empty((a.size // _BPS[intype]) * _BPS[outtype], dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        n = ctypes.c_size_t(0)
        self._check(self._lib.doppler_b200_mix(self._ctx, _ptr(a), a.size, intype, outtype, ctypes.c_float(shift_hz),
                                               int(samplerate), ctypes.byref(sn), _ptr(out), out.size, ctypes.byref(n)))
        return out[:n.value], sn.value

    def mix_blocks(self, inbuf, intype, outtype, shifts_hz, samplerate, samplenum=0, block_bytes=BUFFER_SIZE):
        """One shift per `block_bytes` of input (track mode, main.rs:177)."""
        a = _as_bytes_array(inbuf)
        sh = np.ascontiguousarray(shifts_hz, dtype=np.float32)
        out = np.empty((a.size // _BPS[intype]) * _BPS[outtype], dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        n = ctypes.c_size_t(0)
        self._check(self._lib.doppler_b200_mix_blocks(self._ctx, _ptr(a), a.size, intype, outtype, _ptr(sh), sh.size,
                                                      block_bytes, int(samplerate), ctypes.byref(sn), _ptr(out),
                                                      out.size, ctypes.byref(n)))
        return out[:n.value], sn.value

    # -- device-resident (raw pointers; used by bench.py with torch tensors) -------------
    def mix_dev(self, d_in, in_len, intype, outtype, shift_hz, samplerate, samplenum, d_out, out_cap, stream=None):
        sn = ctypes.c_uint32(samplenum)
        self._check(self._lib.doppler_b200_mix_dev(self._ctx, ctypes.c_void_p(d_in), in_len, intype, outtype,
                                                   ctypes.c_float(shift_hz), int(samplerate), ctypes.byref(sn),
                                                   ctypes.c_void_p(d_out), out_cap, ctypes.c_void_p(stream or 0)))
        return sn.value

    def mix_blocks_dev(self, d_in, in_len, intype, outtype, shifts_hz, samplerate, samplenum, d_out, out_cap,
                       block_bytes=BUFFER_SIZE, stream=None):
        sh = np.ascontiguousarray(shifts_hz, dtype=np.float32)
        sn = ctypes.c_uint32(samplenum)
        self._check(self._lib.doppler_b200_mix_blocks_dev(self._ctx, ctypes.c_void_p(d_in), in_len, intype, outtype,
                                                          _ptr(sh), sh.size, block_bytes, int(samplerate),
                                                          ctypes.byref(sn), ctypes.c_void_p(d_out), out_cap,
                                                          ctypes.c_void_p(stream or 0)))
        return sn.value

    # -- probes --------------------------------------------------------------------------
    def phasor_probe(self, r, n0, count):
        c = np.empty(count, dtype=np.float32)
        s = np.empty(count, dtype=np.float32)
        self._check(self._lib.doppler_b200_phasor_probe(self._ctx, ctypes.c_float(r), int(n0), count, _ptr(c), _ptr(s)))
        return c, s

    def sincosf_probe(self, first_bits, stride, count):
        s = np.empty(count, dtype=np.float32)
        c = np.empty(count, dtype=np.float32)
        self._check(self._lib.doppler_b200_sincosf_probe(self._ctx, int(first_bits), int(stride), count, _ptr(s), _ptr(c)))
        return s, c


class Decimator:
    """Mix + decimate-by-M FIR in one pass (include/doppler_b200.h: doppler_b200_decim_*; not in the reference).
    Carries the FIR history and the stream position across calls; the caller carries samplenum."""

    def __init__(self, mixer, taps, decimation):
        self._lib = mixer._lib
        self._mixer = mixer
        h = np.ascontiguousarray(taps, dtype=np.float32)
        self._d = ctypes.c_void_p()
        self.decimation = int(decimation)
        mixer._check(self._lib.doppler_b200_decim_create(mixer._ctx, _ptr(h), h.size, int(decimation), ctypes.byref(self._d)))

    def close(self):
        if getattr(self, "_d", None):
            self._lib.doppler_b200_decim_destroy(self._d)
            self._d = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._mixer._check(self._lib.doppler_b200_decim_reset(self._d))

    @property
    def position(self):
        return int(self._lib.doppler_b200_decim_position(self._d))

    def mix(self, inbuf, intype, outtype, shift_hz, samplerate, samplenum=0):
        a = _as_bytes_array(inbuf)
        out = np.empty((a.size // _BPS[intype] // self.decimation + 2) * _BPS[outtype], dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        n = ctypes.c_size_t(0)
        self._mixer._check(self._lib.doppler_b200_mix_decimate(self._d, _ptr(a), a.size, intype, outtype, ctypes.c_float(shift_hz),
                                                               int(samplerate), ctypes.byref(sn), _ptr(out), out.size, ctypes.byref(n)))
        return out[:n.value], sn.value

    def mix_blocks(self, inbuf, intype, outtype, shifts_hz, samplerate, samplenum=0, block_bytes=BUFFER_SIZE):
        a = _as_bytes_array(inbuf)
        sh = np.ascontiguousarray(shifts_hz, dtype=np.float32)
        out = np.empty((a.size // _BPS[intype] // self.decimation + 2) * _BPS[outtype], dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        n = ctypes.c_size_t(0)
        self._mixer._check(self._lib.doppler_b200_mix_blocks_decimate(self._d, _ptr(a), a.size, intype, outtype, _ptr(sh), sh.size, block_bytes,
                                                                      int(samplerate), ctypes.byref(sn), _ptr(out), out.size, ctypes.byref(n)))
        return out[:n.value], sn.value

    def mix_dev(self, d_in, in_len, intype, outtype, shift_hz, samplerate, samplenum, d_out, out_cap, stream=None):
        """Device buffers; returns (bytes written, samplenum)."""
        sn = ctypes.c_uint32(samplenum)
        n = ctypes.c_size_t(0)
        self._mixer._check(self._lib.doppler_b200_mix_decimate_dev(self._d, ctypes.c_void_p(d_in), in_len, intype, outtype, ctypes.c_float(shift_hz),
                                                                   int(samplerate), ctypes.byref(sn), ctypes.c_void_p(d_out), out_cap, ctypes.byref(n),
                                                                   ctypes.c_void_p(stream or 0)))
        return n.value, sn.value


class MultiMixer:
    """A group of contexts, one per GPU of the box (include/doppler_b200.h: doppler_b200_multi_create).  One stream is
    cut into contiguous time slices on pump-block boundaries, slice d goes to device d; no collective."""

    def __init__(self, devices=None):
        self._lib = _lib.load()
        self._m = ctypes.c_void_p()
        if devices is None:
            rc = self._lib.doppler_b200_multi_create(None, 0, ctypes.byref(self._m))
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self._lib.doppler_b200_multi_create(arr, len(devices), ctypes.byref(self._m))
        if rc != OK:
            raise DopplerError(rc, self._lib.doppler_b200_last_error(None).decode())

    def close(self):
        if getattr(self, "_m", None):
            self._lib.doppler_b200_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.doppler_b200_multi_size(self._m))

    def _check(self, rc):
        if rc != OK:
            raise DopplerError(rc, self._lib.doppler_b200_multi_last_error(self._m).decode())

    @property
    def launch_count(self):
        return int(self._lib.doppler_b200_multi_launch_count(self._m))

    def synchronize(self):
        self._check(self._lib.doppler_b200_multi_synchronize(self._m))

    def mix(self, inbuf, intype, outtype, shift_hz, samplerate, samplenum=0):
        a = _as_bytes_array(inbuf)
        out = np.empty((a.size // _BPS[intype]) * _BPS[outtype], dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        n = ctypes.c_size_t(0)
        self._check(self._lib.doppler_b200_mix_multi(self._m, _ptr(a), a.size, intype, outtype, ctypes.c_float(shift_hz),
                                                     int(samplerate), ctypes.byref(sn), _ptr(out), out.size, ctypes.byref(n)))
        return out[:n.value], sn.value

    def mix_blocks(self, inbuf, intype, outtype, shifts_hz, samplerate, samplenum=0, block_bytes=BUFFER_SIZE):
        a = _as_bytes_array(inbuf)
        sh = np.ascontiguousarray(shifts_hz, dtype=np.float32)
        out = np.empty((a.size // _BPS[intype]) * _BPS[outtype], dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        n = ctypes.c_size_t(0)
        self._check(self._lib.doppler_b200_mix_blocks_multi(self._m, _ptr(a), a.size, intype, outtype, _ptr(sh), sh.size, block_bytes,
                                                            int(samplerate), ctypes.byref(sn), _ptr(out), out.size, ctypes.byref(n)))
        return out[:n.value], sn.value

    def _slices(self, d_in, in_len, d_out, out_cap):
        k = len(self)
        if not (len(d_in) == len(in_len) == len(d_out) == len(out_cap) == k):
            raise DopplerError(EINVAL, f"need one slice per device ({k})")
        return ((ctypes.c_void_p * k)(*d_in), (ctypes.c_size_t * k)(*in_len), (ctypes.c_void_p * k)(*d_out), (ctypes.c_size_t * k)(*out_cap))

    def mix_dev(self, d_in, in_len, intype, outtype, shift_hz, samplerate, samplenum, d_out, out_cap):
        """Device-resident slices: lists of per-device pointers / byte lengths.  Returns the final samplenum."""
        pi, li, po, lo = self._slices(d_in, in_len, d_out, out_cap)
        sn = ctypes.c_uint32(samplenum)
        self._check(self._lib.doppler_b200_mix_multi_dev(self._m, pi, li, intype, outtype, ctypes.c_float(shift_hz), int(samplerate),
                                                         ctypes.byref(sn), po, lo))
        return sn.value

    def mix_blocks_dev(self, d_in, in_len, intype, outtype, shifts_hz, samplerate, samplenum, d_out, out_cap, block_bytes=BUFFER_SIZE):
        pi, li, po, lo = self._slices(d_in, in_len, d_out, out_cap)
        sh = np.ascontiguousarray(shifts_hz, dtype=np.float32)
        sn = ctypes.c_uint32(samplenum)
        self._check(self._lib.doppler_b200_mix_blocks_multi_dev(self._m, pi, li, intype, outtype, _ptr(sh), sh.size, block_bytes,
                                                                int(samplerate), ctypes.byref(sn), po, lo))
        return sn.value


def libm_compatible():
    """True when this host's libm sincosf is the variant the device reproduces (doppler_b200_libm_compatible)."""
    return bool(_lib.load().doppler_b200_libm_compatible())


def slice_bounds(total_samples, nslices, index, block_samples):
    """[begin, end) of time slice `index` (doppler_b200_slice_bounds)."""
    b, e = ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = _lib.load().doppler_b200_slice_bounds(int(total_samples), int(nslices), int(index), int(block_samples), ctypes.byref(b), ctypes.byref(e))
    if rc != OK:
        raise DopplerError(rc, "bad slice_bounds arguments")
    return b.value, e.value


def slice_seeds(samplenum, shifts_hz, block_samples, samplerate, total_samples, nslices):
    """(begins, seeds), nslices + 1 entries each: first sample of every time slice and the reference's samplenum there
    (doppler_b200_slice_seeds).  One shift value = const mode."""
    sh = np.ascontiguousarray(np.atleast_1d(shifts_hz), dtype=np.float32)
    begins = np.zeros(nslices + 1, dtype=np.uint64)
    seeds = np.zeros(nslices + 1, dtype=np.uint32)
    rc = _lib.load().doppler_b200_slice_seeds(int(samplenum), _ptr(sh), sh.size, int(block_samples), int(samplerate), int(total_samples),
                                              int(nslices), _ptr(begins), _ptr(seeds))
    if rc != OK:
        raise DopplerError(rc, "bad slice_seeds arguments")
    return [int(x) for x in begins], [int(x) for x in seeds]


# ---- host-only analytic samplenum (no GPU needed) ------------------------------------------

def samplenum_advance(samplenum, shift_hz, samplerate, count):
    return int(_lib.load().doppler_b200_samplenum_advance(int(samplenum), ctypes.c_float(shift_hz), int(samplerate), int(count)))


def samplenum_advance_blocks(samplenum, shifts_hz, block_samples, samplerate, count):
    sh = np.ascontiguousarray(shifts_hz, dtype=np.float32)
    return int(_lib.load().doppler_b200_samplenum_advance_blocks(int(samplenum), _ptr(sh), sh.size, int(block_samples),
                                                                 int(samplerate), int(count)))


def plan_trace(samplenum, shifts_hz, block_samples, samplerate, count):
    """(per-sample samplenum sequence, final samplenum, number of closed-form pieces)."""
    sh = np.ascontiguousarray(shifts_hz, dtype=np.float32)
    trace = np.empty(count, dtype=np.uint32)
    sn = ctypes.c_uint32(samplenum)
    npieces = _lib.load().doppler_b200_plan_trace(ctypes.byref(sn), _ptr(sh), sh.size, int(block_samples), int(samplerate),
                                                  int(count), _ptr(trace))
    if npieces < 0:
        raise DopplerError(EINVAL, "bad plan_trace arguments")
    return trace, sn.value, int(npieces)


def plan_tiles_trace(intype, outtype, samplenum, shifts_hz, block_samples, samplerate, count, npipes=148 * 20):
    """Host walk of one kernel launch's work decomposition (segment builder + the kernel's tile iterator).
    Returns (samplenum assigned to every sample, times each sample is covered, tail_begin, stats dict)."""
    sh = np.ascontiguousarray(shifts_hz, dtype=np.float32)
    trace = np.zeros(count, dtype=np.uint32)
    cover = np.zeros(count, dtype=np.uint32)
    stats = np.zeros(8, dtype=np.uint64)
    tb = _lib.load().doppler_b200_plan_tiles_trace(int(intype), int(outtype), int(samplenum), _ptr(sh), sh.size, int(block_samples),
                                                   int(samplerate), int(count), int(npipes), _ptr(trace), _ptr(cover), _ptr(stats))
    if tb < 0:
        raise DopplerError(EINVAL, f"plan_tiles_trace failed ({tb})")
    return trace, cover, int(tb), dict(zip(("segments", "column_segments", "units", "tiles", "column_tiles", "windows", "column_samples", "tile_samples"),
                                         (int(x) for x in stats)))


def doppler_hz(range_rate_km_sec, frequency):
    """main.rs:163."""
    return float(_lib.load().doppler_b200_doppler_hz(float(range_rate_km_sec), int(frequency)))


def replay_schedule(doppler_hz_by_second, offset, samplerate, intype, in_len):
    """The f32 shift of every 8192-byte block as the reference's replay driver forms it (main.rs:155-184)."""
    tab = np.ascontiguousarray(doppler_hz_by_second, dtype=np.float64)
    cap = in_len // BUFFER_SIZE + 1
    out = np.empty(cap, dtype=np.float32)
    n = _lib.load().doppler_b200_replay_schedule(_ptr(tab), tab.size, int(offset), int(samplerate), int(intype), int(in_len),
                                                 _ptr(out), cap)
    if n == 0:
        raise DopplerError(EINVAL, "bad replay_schedule arguments")
    return out[:n]


# ---- module-level functions with the reference's names (default context on cuda:0) ---------
_default = None


def _ctx():
    global _default
    if _default is None:
        _default = Mixer(0)
    return _default


def convert_iqi16_to_complex(inbuf):
    return _ctx().convert_iqi16_to_complex(inbuf)


def convert_iqf32_to_complex(inbuf):
    return _ctx().convert_iqf32_to_complex(inbuf)


def shift_frequency(inbuf, samplenum, shift_hz, samplerate):
    return _ctx().shift_frequency(inbuf, samplenum, shift_hz, samplerate)
