#!/usr/bin/env python
"""tools/ncu_traffic.py -- DRAM traffic of ONE mixer launch at the bench's own launch size, measured by ncu.

Two roles:
  * run under ncu (no --measure): warms up, then issues launches of the bench workload (const f32->i16 @ 10 Msps,
    --shift 100000, `--samples` per launch) for ncu to capture;
  * `measure(samples)`: spawns exactly that under
        ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none
    (the recipe of B200_PROFILING.md), parses the CSV and returns bytes per launch.  bench.py calls it AFTER its timed
    regions (rank 0, one GPU) to fill roofline.traffic; a number taken under ncu is never a bench value -- only
    the byte counts are used.
"""
import argparse
import csv
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KERNEL_REGEX = "mix_grid_kernel"


def workload(samples, launches):
    import torch

    import doppler_b200
    from doppler_b200 import F32, I16
    dev = torch.device("cuda", 0)
    x = torch.empty(2 * samples, dtype=torch.float32, device=dev)
    x.uniform_(-0.7, 0.7)
    y = torch.empty(2 * samples, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    m = doppler_b200.Mixer(0)
    for _ in range(launches):
        m.mix_dev(x.data_ptr(), 8 * samples, F32, I16, 100000.0, 10_000_000, 0, y.data_ptr(), 4 * samples)
        m.synchronize()
    m.close()


def measure(samples, timeout=420, keep_csv=None):
    """-> dict(bytes_read, bytes_write, bytes, duration_ns, samples) for one launch, or raises."""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        raise RuntimeError("ncu not found")
    with tempfile.TemporaryDirectory() as td:
        log = os.path.join(td, "traffic.csv")
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
               "--print-units", "base", "-k", f"regex:{KERNEL_REGEX}", "--launch-skip", "1", "--launch-count", "1", "--csv",
               "--log-file", log, sys.executable, os.path.abspath(__file__), "--samples", str(samples), "--launches", "3"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, cwd=ROOT)
        if r.returncode != 0 or not os.path.exists(log):
            raise RuntimeError(f"ncu failed ({r.returncode}): {r.stdout[-400:]}")
        text = open(log).read()
        if keep_csv:
            os.makedirs(os.path.dirname(keep_csv), exist_ok=True)
            open(keep_csv, "w").write(text)
    lines = text.splitlines()
    start = next(i for i, line in enumerate(lines) if line.startswith('"ID"'))
    vals = {}
    kernel = None
    for row in csv.DictReader(lines[start:]):
        kernel = row.get("Kernel Name", kernel)
        vals[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
    return {"bytes_read": rd, "bytes_write": wr, "bytes": rd + wr, "duration_ns": vals.get("gpu__time_duration.sum"),
            "samples": samples, "kernel": kernel}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=640_000_000)
    ap.add_argument("--launches", type=int, default=3)
    ap.add_argument("--measure", action="store_true")
    ap.add_argument("--csv", default=None)
    a = ap.parse_args()
    if a.measure:
        print(json.dumps(measure(a.samples, keep_csv=a.csv)))
    else:
        workload(a.samples, a.launches)
