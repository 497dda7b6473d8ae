#!/usr/bin/env python
"""tools/sweep.py -- device-resident throughput sweep of the mixer (BASELINE configs[4] / SURVEY 8d cfg5).

For every case: 3 warm-ups, then `iters` launches timed one by one with CUDA events on the
launch stream; L2 is flushed (a 512 MB memset) before each timed launch when the buffers are
smaller than 2x L2.  Reports median and best Msamples/s and the achieved fraction of the
measured HBM copy peak (MEASURED_PEAKS.json).  One JSON object per line + a markdown table.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import doppler_b200  # noqa: E402
from doppler_b200 import F32, I16  # noqa: E402

BPS = {I16: 4, F32: 8}
NAME = {I16: "i16", F32: "f32"}


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run_case(mixer, stream, flush, label, intype, outtype, shift, fs, n, iters):
    dev = torch.device("cuda", 0)
    x = torch.empty(n * BPS[intype], dtype=torch.uint8, device=dev)
    if intype == F32:
        x.view(torch.float32).uniform_(-0.7, 0.7)
    else:
        x.view(torch.int16).copy_(torch.randint(-20000, 20000, (2 * n,), device=dev, dtype=torch.int16))
    y = torch.empty(n * BPS[outtype], dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    need_flush = (x.numel() + y.numel()) < 2 * 126e6
    times = []
    with torch.cuda.stream(stream):
        for i in range(3 + iters):
            if need_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            mixer.mix_dev(x.data_ptr(), x.numel(), intype, outtype, shift, fs, 0, y.data_ptr(), y.numel(), stream=stream.cuda_stream)
            e1.record(stream)
            stream.synchronize()
            if i >= 3:
                times.append(e0.elapsed_time(e1) * 1e-3)
    med, best = statistics.median(times), min(times)
    bps = BPS[intype] + BPS[outtype]
    rec = {"case": label, "in": NAME[intype], "out": NAME[outtype], "shift_hz": shift, "samplerate": fs, "samples": n,
           "median_us": med * 1e6, "best_us": best * 1e6, "msps_median": n / med / 1e6, "msps_best": n / best / 1e6,
           "gbs_median": n * bps / med / 1e9, "frac_of_measured_peak": n * bps / med / 1e9 / peak(), "l2_flushed": need_flush,
           "iters": iters}
    del x, y
    torch.cuda.empty_cache()
    return rec


from tools.workloads import overpass_shifts  # noqa: E402,F401  (the BASELINE configs as code)


def run_track_case(mixer, stream, label, intype, outtype, fs, shifts, n, seed, iters):
    dev = torch.device("cuda", 0)
    x = torch.empty(n * BPS[intype], dtype=torch.uint8, device=dev)
    if intype == F32:
        x.view(torch.float32).uniform_(-0.7, 0.7)
    else:
        v = x.view(torch.int16)
        for k in range(0, v.numel(), 1 << 28):
            v[k:k + (1 << 28)].copy_(torch.randint(-20000, 20000, (min(1 << 28, v.numel() - k),), device=dev, dtype=torch.int16))
    y = torch.empty(n * BPS[outtype], dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    # the host plans every call (runs, pieces, segments) before it launches: time `iters` calls queued back to
    # back, so the planning of call i+1 overlaps the kernels of call i as it does for a streaming caller, and
    # one call alone (planning + kernels, nothing to overlap with)
    import time
    times, single = [], []
    with torch.cuda.stream(stream):
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(iters):
                mixer.mix_blocks_dev(x.data_ptr(), x.numel(), intype, outtype, shifts, fs, seed, y.data_ptr(), y.numel(), stream=stream.cuda_stream)
            e1.record(stream)
            stream.synchronize()
            if rep >= 1:
                times.append(e0.elapsed_time(e1) * 1e-3 / iters)
        for rep in range(3):
            t0 = time.perf_counter()
            mixer.mix_blocks_dev(x.data_ptr(), x.numel(), intype, outtype, shifts, fs, seed, y.data_ptr(), y.numel(), stream=stream.cuda_stream)
            t1 = time.perf_counter()
            stream.synchronize()
            single.append((t1 - t0, time.perf_counter() - t0))
    med, best = statistics.median(times), min(times)
    bps = BPS[intype] + BPS[outtype]
    rec = {"case": label, "in": NAME[intype], "out": NAME[outtype], "shift_hz": "schedule", "samplerate": fs, "samples": n,
           "median_us": med * 1e6, "best_us": best * 1e6, "msps_median": n / med / 1e6, "msps_best": n / best / 1e6,
           "gbs_median": n * bps / med / 1e9, "frac_of_measured_peak": n * bps / med / 1e9 / peak(), "l2_flushed": False,
           "iters": iters, "host_plan_ms": min(a for a, _ in single) * 1e3, "single_call_ms": min(b for _, b in single) * 1e3,
           "note": "median over 3 batches of `iters` calls queued back to back (CUDA events around the batch)"}
    del x, y
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None, help="comma-separated substring filters on the case label")
    args = ap.parse_args()
    mixer = doppler_b200.Mixer(0)
    stream = torch.cuda.Stream()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    cases = []
    pairs = [(I16, I16), (I16, F32), (F32, I16), (F32, F32)]
    big = 256_000_000 if not args.quick else 64_000_000
    # (a) phasor source x type pair, large buffers
    for it, ot in pairs:
        cases.append((f"table-smem P=256 {NAME[it]}->{NAME[ot]}", it, ot, -15000.0, 256000, big))
        cases.append((f"table-L2 P=111145 {NAME[it]}->{NAME[ot]}", it, ot, -9876.54, 1_024_000, big))
        cases.append((f"direct periodic P=4.9M {NAME[it]}->{NAME[ot]}", it, ot, 4_000_000.5, 200_000_000, big))
        cases.append((f"direct linear (no reset) {NAME[it]}->{NAME[ot]}", it, ot, 1.0, 2_000_000_000, big))
    # (b) cfg5: const i16->i16, 1 s buffers, r = -15/256 at every rate
    for fs in [256_000, 1_024_000, 10_000_000, 100_000_000, 1_000_000_000, 2_000_000_000]:
        if args.quick and fs > 100_000_000:
            continue
        cases.append((f"cfg5 1s @ {fs} sps i16->i16", I16, I16, -15000.0 * fs / 256000.0, fs, fs))
    cases.append(("cfg5 irregular 7321 Hz @ 1.024 Msps, 1 s", I16, I16, 7321.0, 1_024_000, 1_024_000))
    cases.append(("cfg2 f32->i16 10 Msps shift 100000, 64 s", F32, I16, 100000.0, 10_000_000, 640_000_000 if not args.quick else 64_000_000))
    if args.only:
        cases = [c for c in cases if any(f.replace("_", " ") in c[0] for f in args.only.split(","))]   # "_" stands for a blank (shell-friendly)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    recs = []
    with open(args.out, "w") as f:
        for c in cases:
            rec = run_case(mixer, stream, flush, *c, args.iters)
            recs.append(rec)
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(f"{rec['case']:55s} {rec['msps_median']:12.1f} Msps  {rec['gbs_median']:8.1f} GB/s  {rec['frac_of_measured_peak']:.3f}")
        # (c) track mode at full size: cfg3 (600 s @ 1.024 Msps i16) and one GPU's slice of cfg4 (f32 @ 200 Msps)
        track = []
        if not args.quick:
            from doppler_b200 import slicing
            n3 = 600 * 1_024_000
            track.append(("cfg3 track replay 600 s @ 1.024 Msps i16->i16", I16, I16, 1_024_000,
                          overpass_shifts(1_024_000, 600, 437_505_000, 300.0, 5000, I16, n3), n3, 0))
            total = 60 * 200_000_000
            b, e = slicing.slice_bounds(total, 8, 3, F32)
            sh = overpass_shifts(200_000_000, 60, 4_200_000_000, 30.0, 0, F32, total)
            track.append(("cfg4 track f32->f32 @ 200 Msps, slice 3 of 8 (1.5 G samples)", F32, F32, 200_000_000,
                          sh[b // slicing.block_samples(F32):], e - b, slicing.seed_blocks(sh, F32, 200_000_000, b)))
        for c in track:
            if args.only and not any(f.replace("_", " ") in c[0] for f in args.only.split(",")):
                continue
            rec = run_track_case(mixer, stream, *c, max(3, args.iters // 4))
            recs.append(rec)
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(f"{rec['case']:55s} {rec['msps_median']:12.1f} Msps  {rec['gbs_median']:8.1f} GB/s  {rec['frac_of_measured_peak']:.3f}")
    md = args.out.replace(".jsonl", ".md")
    with open(md, "w") as f:
        f.write("| case | samples | median us | Msamples/s (median) | Msamples/s (best) | GB/s | frac of measured HBM peak | L2 flushed |\n|---|---|---|---|---|---|---|---|\n")
        for r in recs:
            f.write(f"| {r['case']} | {r['samples']} | {r['median_us']:.1f} | {r['msps_median']:.0f} | {r['msps_best']:.0f} | {r['gbs_median']:.0f} | {r['frac_of_measured_peak']:.3f} | {r['l2_flushed']} |\n")
    mixer.close()


if __name__ == "__main__":
    main()
