#!/usr/bin/env python
"""tools/track_scaling.py -- BASELINE configs[3]: track mode, f32 IQ @ 200 Msps, 60 s synthetic overpass (12 G samples),
time-sliced across the ranks of one box (torchrun, one rank per GPU).

Every rank mixes its contiguous slice (whole 8192-byte blocks) with the per-block shift schedule of the reference's
replay driver and the analytically carried samplenum -- no collective on the data path.  Timing: batches of calls
queued back to back, CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks.  Each
rank also checks two windows of its slice against the oracle.  Rank 0 prints one JSON line."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import doppler_b200  # noqa: E402
from doppler_b200 import F32, slicing  # noqa: E402
from tools.workloads import overpass_shifts  # noqa: E402
from tests.oracle_lib import Oracle, same_bits_f32  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    fs, secs = 200_000_000, 60
    total = secs * fs
    nslices = max(world, 8)                                  # the job is cut into 8 slices; with fewer GPUs each takes slice `rank`
    b, e = slicing.slice_bounds(total, nslices, rank, F32)
    shifts = overpass_shifts(fs, secs, 4_200_000_000, 30.0, 0, F32, total)
    bs = slicing.block_samples(F32)
    sl = shifts[b // bs:]
    seed = slicing.seed_blocks(shifts, F32, fs, b)
    n = e - b
    x = torch.empty(n * 8, dtype=torch.uint8, device=dev)
    x.view(torch.float32).uniform_(-0.7, 0.7)
    y = torch.empty(n * 8, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    mixer = doppler_b200.Mixer(local)
    stream = torch.cuda.Stream(device=dev)

    def call():
        return mixer.mix_blocks_dev(x.data_ptr(), x.numel(), F32, F32, sl, fs, seed, y.data_ptr(), y.numel(), stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        call()
    barrier()
    iters, times = 5, []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(iters):
            call()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        times.append(ms)
    # parity spot check on this rank's slice
    oracle = Oracle()
    ok = 1
    for w0 in (0, (n // 2) // bs * bs):
        w1 = min(w0 + 65536, n)
        want, _ = oracle.mix_blocks(x[w0 * 8:w1 * 8].cpu().numpy(), F32, F32, sl[w0 // bs:], fs, samplenum=slicing.seed_blocks(shifts, F32, fs, b + w0))
        ok &= int(same_bits_f32(y[w0 * 8:w1 * 8].cpu().numpy(), want))
    if world > 1:
        t = torch.tensor([ok], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = int(t.item())
    if rank == 0:
        ms = statistics.median(times)
        peak = 6650.0
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        print(json.dumps({"workload": "cfg4: track f32->f32 @ 200 Msps, 60 s overpass, 8 time slices", "n_gpus": world,
                          "samples_per_gpu": n, "ms_per_call": ms, "msps_total": world * n / ms / 1e3,
                          "gbs_per_gpu": n * 16 / ms / 1e6, "frac_of_measured_peak_per_gpu": n * 16 / ms / 1e6 / peak,
                          "parity_windows_ok": bool(ok), "collective_on_data_path": False}))
    mixer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
