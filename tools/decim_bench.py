#!/usr/bin/env python
"""tools/decim_bench.py -- the fused mix + decimating FIR stage, device-resident: register-blocked kernel
against the generic kernel, interleaved in one process, over type pairs and filters.  Each case is checked bit for bit against
the oracle's specification on its first 2^18 inputs.  JSON lines: input Msamples/s, fraction of the stage's HBM roofline
(input bytes + output bytes / M per input sample against the measured copy peak)."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import doppler_b200  # noqa: E402
from doppler_b200 import F32, I16  # noqa: E402
from tests.oracle_lib import Oracle  # noqa: E402

BPS = {I16: 4, F32: 8}
NAME = {I16: "i16", F32: "f32"}


def lowpass(ntaps, cutoff):
    t = np.arange(ntaps) - (ntaps - 1) / 2.0
    h = np.sinc(2 * cutoff * t) * np.hamming(ntaps)
    return (h / h.sum()).astype(np.float32)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "decim_bench.jsonl")
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256_000_000
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream()
    mixer = doppler_b200.Mixer(0)
    oracle = Oracle()
    #        intype outtype fs          shift      M   ntaps
    cases = [(F32, I16, 10_000_000, 100000.0, 8, 49),      # bench.py's f4 configuration
             (I16, I16, 1_024_000, 5000.0, 8, 49),
             (F32, F32, 10_000_000, 100000.0, 8, 49),
             (I16, F32, 256_000, -15000.0, 4, 33),
             (F32, I16, 10_000_000, 100000.0, 2, 15),
             (F32, I16, 10_000_000, 100000.0, 16, 97),
             (F32, I16, 10_000_000, 100000.0, 5, 41),
             (I16, I16, 256_000, -15000.0, 32, 128),
             (F32, I16, 1_024_000, -9876.54, 8, 49)]       # long period: no table, per-sample evaluation in phase A
    if os.environ.get("DECIM_CASES"):                              # e.g. DECIM_CASES=0,1 (profiling one case under ncu)
        cases = [cases[int(i)] for i in os.environ["DECIM_CASES"].split(",")]
    with open(out, "w") as f:
        for it, ot, fs, shift, M, ntaps in cases:
            taps = lowpass(ntaps, 0.4 / M)
            dec = doppler_b200.Decimator(mixer, taps, M)
            x = torch.empty(n * BPS[it], dtype=torch.uint8, device=dev)
            if it == F32:
                x.view(torch.float32).uniform_(-0.7, 0.7)
            else:
                x.view(torch.int16).random_(-32768, 32768)
            y = torch.empty((n // M + 2) * BPS[ot], dtype=torch.uint8, device=dev)
            w = 1 << 18
            want, _ = oracle.mix_decimate(x[:w * BPS[it]].cpu().numpy(), it, ot, shift, fs, taps, M)
            variants = [("register-blocked", 0, None), ("generic", 1, None)]
            if os.environ.get("DECIM_SLOTS"):                      # stage-size sweep of the register-blocked kernel
                variants = [(f"register-blocked slots={v}", 0, int(v)) for v in os.environ["DECIM_SLOTS"].split(",")]
            res, ok = {v[0]: [] for v in variants}, {}
            for rep in range(3):
                for name, variant, slots in variants:
                    if name == "generic" and rep > 0 and n > 64_000_000:
                        continue                                   # the generic kernel is 4x slower: one repeat
                    mixer.tune(decim_variant=variant, decim_stage_slots=slots)
                    dec.reset()
                    y.zero_()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    dec.mix_dev(x.data_ptr(), x.numel(), it, ot, shift, fs, 0, y.data_ptr(), y.numel(), stream=stream.cuda_stream)
                    e1.record(stream)
                    stream.synchronize()
                    res[name].append(e0.elapsed_time(e1) * 1e-3)
                    ok[name] = bool(np.array_equal(y[:want.size].cpu().numpy(), want))
            bps = BPS[it] + BPS[ot] / M
            row = {"case": f"{NAME[it]}->{NAME[ot]} fs={fs} shift={shift} M={M} ntaps={ntaps}", "samples": n, "bytes_per_input_sample": bps}
            for name, tt in res.items():
                t = statistics.median(tt)
                row[name] = {"ms": t * 1e3, "msps_in": n / t / 1e6, "frac": n * bps / t / 1e9 / peak, "parity_ok": ok[name]}
            print(json.dumps(row), flush=True)
            f.write(json.dumps(row) + "\n")
            dec.close()
            del x, y
            torch.cuda.empty_cache()
    mixer.tune(decim_variant=0)
    mixer.close()


if __name__ == "__main__":
    main()
