#!/usr/bin/env python
"""tools/latency.py -- small-launch and per-call latency of the mixer (BASELINE configs[4]'s small end; the reference's
own operating point, README.md:53 / main.rs:49).

(a) device-resident launches, const i16->i16 r = -15/256, sizes 2 k .. 100 M samples: the latency-shaped small kernel
    against the persistent bulk-async kernel (doppler_b200_tune SMALL_MAX_SAMPLES), each launch timed alone with CUDA
    events after an L2 flush -- picks the threshold;
(b) host-buffer calls at the reference's granularity (one 8192-byte block per call) and a few larger sizes: zero-copy
    tiny path against the staged pipeline (TINY_HOST_BYTES), pageable and pinned caller buffers, wall clock per call.
One JSON object per line."""
import argparse
import ctypes
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import doppler_b200  # noqa: E402
from doppler_b200 import F32, I16, _lib  # noqa: E402


def device_side(out):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    mixers = {"small": doppler_b200.Mixer(0), "bulk": doppler_b200.Mixer(0)}
    mixers["small"].tune(small_max_samples=1 << 30)
    mixers["bulk"].tune(small_max_samples=0)
    # the floor of this protocol: the smallest possible kernel (a 4-byte fill) between the same two events after the same flush
    one = torch.zeros(1, dtype=torch.int32, device=dev)
    ts = []
    with torch.cuda.stream(stream):
        for i in range(23):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            one.fill_(i)
            e1.record(stream)
            stream.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
    out({"what": "protocol floor", "kernel": "torch fill_ of 4 bytes", "us_median": statistics.median(ts), "us_best": min(ts)})
    cases = [("P=256 table", -15000.0, 256000), ("irregular direct", 7321.0, 1_024_000)]
    for label, shift, fs in cases:
        for n in [2048, 65_536, 256_000, 1_024_000, 2_000_000, 4_000_000, 10_000_000, 20_000_000, 40_000_000, 100_000_000]:
            x = torch.randint(-20000, 20000, (2 * n,), device=dev, dtype=torch.int16)
            y = torch.empty(2 * n, dtype=torch.int16, device=dev)
            torch.cuda.synchronize()
            rec = {"what": "device launch", "case": label, "samples": n}
            for name, m in mixers.items():
                ts = []
                with torch.cuda.stream(stream):
                    for i in range(3 + 20):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        m.mix_dev(x.data_ptr(), 4 * n, I16, I16, shift, fs, 0, y.data_ptr(), 4 * n, stream=stream.cuda_stream)
                        e1.record(stream)
                        stream.synchronize()
                        if i >= 3:
                            ts.append(e0.elapsed_time(e1) * 1e3)
                rec[name + "_us_median"] = statistics.median(ts)
                rec[name + "_us_best"] = min(ts)
            # back-to-back (no flush, no sync between launches): what a streaming caller sees per launch
            for name, m in mixers.items():
                with torch.cuda.stream(stream):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(50):
                        m.mix_dev(x.data_ptr(), 4 * n, I16, I16, shift, fs, 0, y.data_ptr(), 4 * n, stream=stream.cuda_stream)
                    e1.record(stream)
                    stream.synchronize()
                rec[name + "_us_back_to_back"] = e0.elapsed_time(e1) * 1e3 / 50
            out(rec)
            del x, y
    for m in mixers.values():
        m.close()


def host_side(out):
    lib = _lib.load()
    for typ, tname, bps in ((I16, "i16", 4), (F32, "f32", 8)):
        for nbytes in (8192, 65536, 131072, 1 << 20):
            n = nbytes // bps
            for pinned in (False, True):
                if pinned:
                    hin, hout = lib.doppler_b200_host_alloc(nbytes), lib.doppler_b200_host_alloc(nbytes)
                    src = np.ctypeslib.as_array(ctypes.cast(hin, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,))
                else:
                    a_in, a_out = np.zeros(nbytes, dtype=np.uint8), np.zeros(nbytes, dtype=np.uint8)
                    hin, hout, src = a_in.ctypes.data, a_out.ctypes.data, a_in
                src[:] = np.random.default_rng(1).integers(0, 64, nbytes, dtype=np.uint8)
                rec = {"what": "host call doppler_b200_mix", "type": tname, "bytes_per_call": nbytes, "samples": n, "caller_buffers": "pinned" if pinned else "pageable"}
                for name, tiny in (("zero_copy", 8 << 20), ("staged", 0)):
                    m = doppler_b200.Mixer(0)
                    m.tune(tiny_host_bytes=tiny)
                    sn = ctypes.c_uint32(0)
                    got = ctypes.c_size_t(0)
                    iters = 3000 if nbytes <= 131072 else 500

                    def call():
                        return lib.doppler_b200_mix(m._ctx, hin, nbytes, typ, typ, ctypes.c_float(5000.0), 1_024_000, ctypes.byref(sn), hout, nbytes, ctypes.byref(got))
                    for _ in range(50):
                        assert call() == 0
                    t0 = time.perf_counter()
                    for _ in range(iters):
                        call()
                    rec[name + "_us_per_call"] = (time.perf_counter() - t0) / iters * 1e6
                    m.close()
                out(rec)
                if pinned:
                    lib.doppler_b200_host_free(hin)
                    lib.doppler_b200_host_free(hout)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "latency.jsonl"))
    ap.add_argument("--only", default=None, choices=[None, "device", "host"])
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        def out(rec):
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec))
        if a.only in (None, "device"):
            device_side(out)
        if a.only in (None, "host"):
            host_side(out)
