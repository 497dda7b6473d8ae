#!/usr/bin/env python
"""tools/ab_seg.py -- A/B of the segmented kernel's variants IN ONE PROCESS, interleaved (boxes differ by +-10 % under
sustained load, so only same-session comparisons count): 4-warp pipelines vs per-warp pipelines (i16->i16), guided chunk
claims vs single-unit claims.  Cases: COLUMN at 256 M samples for all type pairs, cfg3, one cfg4 slice.  JSON lines."""
import itertools
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import doppler_b200  # noqa: E402
from doppler_b200 import F32, I16, slicing  # noqa: E402
from tools import workloads as W  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ab_seg.jsonl")
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream()
    variants = {}
    for sv, mc in itertools.product((0, 1), (8, 1)):
        m = doppler_b200.Mixer(0)
        m.tune(seg_variant=sv, max_claim=mc)
        variants[f"{'4-warp pipelines' if sv else 'per-warp pipelines'} claim<={mc}"] = m
    cases = []
    n = 256_000_000
    for it, ot in ((I16, I16), (I16, F32), (F32, I16), (F32, F32)):
        cases.append((f"COLUMN P=111145 {W.NAME[it]}->{W.NAME[ot]}", it, ot, 1_024_000, None, -9876.54, n, 0))
    c3 = W.CFG3
    cases.append(("cfg3", I16, I16, c3["fs"], W.cfg_schedule(c3), None, c3["secs"] * c3["fs"], 0))
    c4 = W.CFG4
    total = c4["secs"] * c4["fs"]
    sh4 = W.cfg_schedule(c4)
    b, e = slicing.slice_bounds(total, 8, 3, F32)
    cases.append(("cfg4 slice 3/8", F32, F32, c4["fs"], sh4[b // 1024:], None, e - b, slicing.seed_blocks(sh4, F32, c4["fs"], b)))
    with open(out, "w") as f:
        for label, it, ot, fs, shifts, shift, n, seed in cases:
            x = torch.empty(n * W.BPS[it], dtype=torch.uint8, device=dev)
            W.fill_device(x, it)
            y = torch.empty(n * W.BPS[ot], dtype=torch.uint8, device=dev)
            res = {k: [] for k in variants}
            for rep in range(4):
                for name, m in variants.items():
                    def call():
                        if shifts is None:
                            m.mix_dev(x.data_ptr(), x.numel(), it, ot, shift, fs, seed, y.data_ptr(), y.numel(), stream=stream.cuda_stream)
                        else:
                            m.mix_blocks_dev(x.data_ptr(), x.numel(), it, ot, shifts, fs, seed, y.data_ptr(), y.numel(), stream=stream.cuda_stream)
                    call()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(5):
                        call()
                    e1.record(stream)
                    stream.synchronize()
                    if rep >= 1:
                        res[name].append(e0.elapsed_time(e1) / 5)
            bps = W.BPS[it] + W.BPS[ot]
            rec = {"case": label, "samples": n}
            for name, ts in res.items():
                ms = statistics.median(ts)
                rec[name] = {"ms": ms, "frac": n * bps / ms / 1e6 / peak}
            f.write(json.dumps(rec) + "\n")
            print(label, {k: round(v["frac"], 3) for k, v in rec.items() if isinstance(v, dict)})
            del x, y
            torch.cuda.empty_cache()
    for m in variants.values():
        m.close()


if __name__ == "__main__":
    main()
