"""ctypes harness around oracle/liboracle.so (the CPU restatement of the reference) and
oracle/_ref/libcomplex_ref.so (the reference's own complex.c, compiled unmodified).

Test infrastructure only.  Nothing under doppler_b200/ imports this."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcomplex_ref.so")

I16, F32 = 0, 1
BPS = {I16: 4, F32: 8}
BUFFER_SIZE = 8192


class C32(ctypes.Structure):
    _fields_ = [("re", ctypes.c_float), ("im", ctypes.c_float)]


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a.size else ctypes.c_void_p(0)


class Oracle:
    def __init__(self, use_ref=True):
        self.lib = ctypes.CDLL(ORACLE_SO)
        L = self.lib
        L.oracle_convert_iqi16_to_complex.restype = ctypes.c_long
        L.oracle_convert_iqf32_to_complex.restype = ctypes.c_long
        L.oracle_mix.restype = ctypes.c_long
        L.oracle_const_stream.restype = ctypes.c_long
        L.oracle_track_replay_stream.restype = ctypes.c_long
        L.oracle_mix_blocks.restype = ctypes.c_long
        L.oracle_samplenum_advance.restype = ctypes.c_uint32
        L.oracle_samplenum_advance.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_uint32, ctypes.c_uint64]
        L.oracle_samplenum_trace.restype = ctypes.c_uint32
        L.oracle_samplenum_trace.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p]
        L.oracle_doppler_hz.restype = ctypes.c_double
        L.oracle_doppler_hz.argtypes = [ctypes.c_double, ctypes.c_uint32]
        L.oracle_bench_const.restype = ctypes.c_double
        L.oracle_bench_blocks.restype = ctypes.c_double
        L.oracle_theta.restype = ctypes.c_float
        L.oracle_theta.argtypes = [ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32]
        L.oracle_libc_version.restype = ctypes.c_char_p
        L.oracle_set_ccexpf.argtypes = [ctypes.c_void_p]
        self.ref = None
        if os.path.exists(REF_SO):
            self.ref = ctypes.CDLL(REF_SO)
        self.has_ref = self.ref is not None
        self.use_ref(use_ref)

    # -- which ccexpf the loop calls -----------------------------------------------------
    def use_ref(self, on):
        """True: the reference's compiled complex.c; False: the oracle's restatement."""
        if on and self.ref is not None:
            self.lib.oracle_set_ccexpf(ctypes.cast(self.ref.ccexpf, ctypes.c_void_p))
            self.using_ref = True
        else:
            self.lib.oracle_set_ccexpf(None)
            self.using_ref = False

    def libc_version(self):
        return self.lib.oracle_libc_version().decode()

    # -- pieces ---------------------------------------------------------------------------
    def ccexpf(self, re, im):
        z = C32(re, im)
        self.lib.oracle_ccexpf(ctypes.byref(z))
        return z.re, z.im

    def ccexpf_restated(self, re, im):
        was = self.using_ref
        self.use_ref(False)
        out = self.ccexpf(re, im)
        self.use_ref(was)
        return out

    def ccexpf_reference(self, re, im):
        z = C32(re, im)
        self.ref.ccexpf(ctypes.byref(z))
        return z.re, z.im

    def convert_iqi16_to_complex(self, buf):
        a = np.frombuffer(bytes(buf), dtype=np.uint8)
        out = np.empty(a.size // 4 + 1, dtype=np.complex64)
        n = self.lib.oracle_convert_iqi16_to_complex(_p(a), ctypes.c_size_t(a.size), _p(out))
        return None if n < 0 else out[:n].copy()

    def convert_iqf32_to_complex(self, buf):
        a = np.frombuffer(bytes(buf), dtype=np.uint8)
        out = np.empty(a.size // 8 + 1, dtype=np.complex64)
        n = self.lib.oracle_convert_iqf32_to_complex(_p(a), ctypes.c_size_t(a.size), _p(out))
        return None if n < 0 else out[:n].copy()

    def shift_frequency(self, inbuf, samplenum, shift_hz, samplerate):
        a = np.ascontiguousarray(inbuf, dtype=np.complex64)
        out = np.empty_like(a)
        sn = ctypes.c_uint32(samplenum)
        self.lib.oracle_shift_frequency(_p(a), ctypes.c_size_t(a.size), ctypes.byref(sn), ctypes.c_float(shift_hz),
                                        ctypes.c_uint32(samplerate), _p(out))
        return out, sn.value

    def mix(self, buf, intype, outtype, shift_hz, samplerate, samplenum=0):
        a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1))
        out = np.empty((a.size // BPS[intype]) * BPS[outtype] + 8, dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        n = self.lib.oracle_mix(_p(a), ctypes.c_size_t(a.size), intype, outtype, ctypes.c_float(shift_hz),
                                ctypes.c_uint32(samplerate), ctypes.byref(sn), _p(out))
        if n < 0:
            return None, samplenum
        return out[:n].copy(), sn.value

    def mix_blocks(self, buf, intype, outtype, shifts, samplerate, samplenum=0):
        a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1))
        sh = np.ascontiguousarray(shifts, dtype=np.float32)
        out = np.empty((a.size // BPS[intype]) * BPS[outtype] + 8, dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        n = self.lib.oracle_mix_blocks(_p(a), ctypes.c_size_t(a.size), intype, outtype, _p(sh), ctypes.c_size_t(sh.size),
                                       ctypes.c_uint32(samplerate), ctypes.byref(sn), _p(out))
        if n < 0:
            return None, samplenum
        return out[:n].copy(), sn.value

    def const_stream(self, buf, intype, outtype, shift, samplerate):
        """main.rs const driver over an in-memory stdin.  Returns (stdout bytes, final samplenum, panicked)."""
        a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1))
        out = np.empty((a.size // BPS[intype] + 1) * BPS[outtype] + 8, dtype=np.uint8)
        sn = ctypes.c_uint32(0)
        n = self.lib.oracle_const_stream(_p(a), ctypes.c_size_t(a.size), intype, outtype, ctypes.c_int32(shift),
                                         ctypes.c_uint32(samplerate), _p(out), ctypes.byref(sn))
        if n < 0:
            return out[:(-n - 1)].copy(), None, True
        return out[:n].copy(), sn.value, False

    def track_replay_stream(self, buf, intype, outtype, doppler_hz_by_second, offset, samplerate):
        """main.rs replay driver.  Returns (stdout bytes, final samplenum, per-block f32 shifts, panicked)."""
        a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1))
        tab = np.ascontiguousarray(doppler_hz_by_second, dtype=np.float64)
        out = np.empty((a.size // BPS[intype] + 1) * BPS[outtype] + 8, dtype=np.uint8)
        cap = a.size // BUFFER_SIZE + 2
        shifts = np.zeros(cap, dtype=np.float32)
        sn = ctypes.c_uint32(0)
        nb = ctypes.c_size_t(0)
        n = self.lib.oracle_track_replay_stream(_p(a), ctypes.c_size_t(a.size), intype, outtype, _p(tab),
                                                ctypes.c_size_t(tab.size), ctypes.c_int32(offset), ctypes.c_uint32(samplerate),
                                                _p(out), ctypes.byref(sn), _p(shifts), ctypes.c_size_t(cap), ctypes.byref(nb))
        if n < 0:
            return out[:(-n - 1)].copy(), None, shifts[:nb.value].copy(), True
        return out[:n].copy(), sn.value, shifts[:nb.value].copy(), False

    def samplenum_advance(self, samplenum, shift_hz, samplerate, count):
        return int(self.lib.oracle_samplenum_advance(samplenum, shift_hz, samplerate, count))

    def samplenum_trace(self, samplenum, shift_hz, samplerate, count):
        tr = np.empty(count, dtype=np.uint32)
        sn = self.lib.oracle_samplenum_trace(samplenum, shift_hz, samplerate, count, _p(tr))
        return tr, int(sn)

    def doppler_hz(self, range_rate_km_sec, frequency):
        return float(self.lib.oracle_doppler_hz(range_rate_km_sec, frequency))

    def theta(self, shift_hz, samplerate, n):
        return float(self.lib.oracle_theta(shift_hz, samplerate, n))

    def sincosf_batch(self, theta):
        t = np.ascontiguousarray(theta, dtype=np.float32)
        s = np.empty_like(t)
        c = np.empty_like(t)
        self.lib.oracle_sincosf_batch(_p(t), ctypes.c_size_t(t.size), _p(s), _p(c))
        return s, c

    def bench_const(self, buf, nsamples, intype, outtype, shift_hz, samplerate, threads):
        a = np.ascontiguousarray(buf.view(np.uint8).reshape(-1))
        out = np.empty(nsamples * BPS[outtype], dtype=np.uint8)
        t = self.lib.oracle_bench_const(_p(a), ctypes.c_size_t(nsamples), intype, outtype, ctypes.c_float(shift_hz),
                                        ctypes.c_uint32(samplerate), _p(out), ctypes.c_int(threads))
        return float(t), out


    def bench_blocks(self, buf, nsamples, intype, outtype, shifts, samplerate, threads, samplenum=0):
        """Threaded track-mode mix over a per-block schedule.  Returns (seconds, output bytes, final samplenum)."""
        a = np.ascontiguousarray(buf.view(np.uint8).reshape(-1))
        sh = np.ascontiguousarray(shifts, dtype=np.float32)
        out = np.empty(nsamples * BPS[outtype], dtype=np.uint8)
        sn = ctypes.c_uint32(samplenum)
        t = self.lib.oracle_bench_blocks(_p(a), ctypes.c_size_t(nsamples), intype, outtype, _p(sh), ctypes.c_size_t(sh.size),
                                         ctypes.c_uint32(samplerate), ctypes.byref(sn), _p(out), ctypes.c_int(threads))
        if t < 0:
            raise RuntimeError("oracle_bench_blocks: schedule shorter than the stream")
        return float(t), out, sn.value

    def mix_decimate(self, buf, intype, outtype, shifts, samplerate, taps, M, state=None):
        """The fused mix + decimating FIR specification (oracle_mix_decimate).  `shifts`: one per 8192-byte block (a scalar
        is repeated).  `state` = dict(samplenum, hist, pos) carried across calls (None: a fresh stream).
        Returns (output bytes, state)."""
        a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1))
        h = np.ascontiguousarray(taps, dtype=np.float32)
        n = a.size // BPS[intype]
        nblocks = max(1, (a.size + BUFFER_SIZE - 1) // BUFFER_SIZE)
        sh = np.ascontiguousarray(np.broadcast_to(np.asarray(shifts, dtype=np.float32), (nblocks,)) if np.ndim(shifts) == 0 else shifts, dtype=np.float32)
        if state is None:
            state = {"samplenum": 0, "hist": np.zeros(2 * max(h.size - 1, 1), dtype=np.float32), "pos": 0}
        hist = np.ascontiguousarray(state["hist"], dtype=np.float32).copy()
        sn = ctypes.c_uint32(state["samplenum"])
        pos = ctypes.c_uint64(state["pos"])
        out = np.empty((n // M + 2) * BPS[outtype], dtype=np.uint8)
        self.lib.oracle_mix_decimate.restype = ctypes.c_long
        w = self.lib.oracle_mix_decimate(_p(a), ctypes.c_size_t(a.size), intype, outtype, _p(sh), ctypes.c_size_t(sh.size),
                                         ctypes.c_uint32(samplerate), ctypes.byref(sn), _p(h), ctypes.c_uint32(h.size), ctypes.c_uint32(M),
                                         _p(hist), ctypes.byref(pos), _p(out))
        if w < 0:
            raise RuntimeError("oracle_mix_decimate: bad arguments")
        return out[:w].copy(), {"samplenum": sn.value, "hist": hist, "pos": pos.value}

    def mix_blocks_threads(self, buf, intype, outtype, shifts, samplerate, samplenum=0, threads=None):
        """oracle.mix_blocks on all host cores (same bytes, same final samplenum)."""
        import os
        a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1))
        if a.size % BPS[intype]:
            return self.mix_blocks(a, intype, outtype, shifts, samplerate, samplenum)
        threads = threads or len(os.sched_getaffinity(0))
        _, out, sn = self.bench_blocks(a, a.size // BPS[intype], intype, outtype, shifts, samplerate, threads, samplenum)
        return out, sn


def same_bits_f32(a, b):
    """Bit equality of two float32 byte streams, NaNs compared as NaN (payload/sign ignored)."""
    fa = np.ascontiguousarray(a).view(np.float32)
    fb = np.ascontiguousarray(b).view(np.float32)
    if fa.shape != fb.shape:
        return False
    ua, ub = fa.view(np.uint32), fb.view(np.uint32)
    nan = np.isnan(fa) & np.isnan(fb)
    return bool(np.all((ua == ub) | nan))
