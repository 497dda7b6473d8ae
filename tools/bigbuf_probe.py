#!/usr/bin/env python
"""tools/bigbuf_probe.py -- why do 4-8 GB buffers run at 0.71-0.75 of peak when 1 GB ones reach 0.96?

Times the table-smem i16->i16 mixer (r = -15/256, P = 256) over buffer sizes and over the distance
between the input and output allocations (both carved from ONE big allocation so the distance is
controlled), with CUDA events on the launch stream.  One JSON line per point.
"""
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


def main():
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    mixer = doppler_b200.Mixer(0)
    stream = torch.cuda.Stream()
    pool = torch.empty(20 << 30, dtype=torch.uint8, device="cuda")
    pool.view(torch.int16).random_(-20000, 20000)
    torch.cuda.synchronize()
    base = pool.data_ptr()
    out = open(os.path.join(ROOT, "gpurun_out", "bigbuf_probe.jsonl"), "w")
    for it, ot in [(I16, I16), (F32, I16)]:
        for n in [128_000_000, 256_000_000, 384_000_000, 512_000_000, 768_000_000, 1_000_000_000]:
            for gap in [0, 1 << 20, (1 << 21) + 4096, 768 << 20]:
                in_bytes, out_bytes = n * BPS[it], n * BPS[ot]
                off_out = (in_bytes + gap + 255) // 256 * 256
                if off_out + out_bytes > pool.numel():
                    continue
                times = []
                with torch.cuda.stream(stream):
                    for i in range(3 + 7):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        mixer.mix_dev(base, in_bytes, it, ot, -15000.0, 256000, 0, base + off_out, out_bytes, stream=stream.cuda_stream)
                        e1.record(stream)
                        stream.synchronize()
                        if i >= 3:
                            times.append(e0.elapsed_time(e1) * 1e-3)
                med = statistics.median(times)
                rec = {"in": it, "out": ot, "samples": n, "gap": gap, "median_us": med * 1e6, "best_us": min(times) * 1e6,
                       "gbs": n * (BPS[it] + BPS[ot]) / med / 1e9, "frac": n * (BPS[it] + BPS[ot]) / med / 1e9 / peak}
                out.write(json.dumps(rec) + "\n")
                out.flush()
                print(rec)
    mixer.close()


if __name__ == "__main__":
    main()
