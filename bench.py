#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the NCO mixer on N B200s (BASELINE.json's metric).

A "step" is one pass of the hot path over one batch of synthetic IQ.  Headline workload (configs[1] of
BASELINE.json): const mode, f32 -> i16, fs = 10 Msps, --shift 100000; the batch is 64 s of that
stream per GPU (640 M complex samples: 5.12 GB in, 2.56 GB out -- far larger than the 126 MB L2,
so no flush is needed between iterations).  With N > 1 (torchrun, one rank per GPU) the stream is
N x 64 s long, cut into contiguous time slices; every rank seeds its slice with the analytically
carried samplenum (doppler_b200_slice_seeds) -- no collective on the data path ("weak").

value      = samples all ranks processed / max-over-ranks device time, inputs resident in HBM
e2e        = same metric through the host-buffer C-ABI call (pinned host in/out, H2D + kernel +
             D2H inside the timed region); e2e.ceiling = the same pipeline with the kernel skipped
             (doppler_b200_pipeline_probe), e2e.pageable = the same call on ordinary pageable memory
roofline   = algorithmic bytes (12 B/sample for f32->i16) / CUDA-event time of the mixer launch,
             against MEASURED_PEAKS.json hbm_gbs; .sustained = the same launches back to back for >= 2 s;
             .traffic = ncu dram bytes of one launch at this launch size (tools/ncu_traffic.py, after timing)
configs    = every other BASELINE config at its full size, after the headline region:
             cfg1 (CLI, 256 k samples), cfg3 (track replay, 614 M samples), cfg4 (track f32 @ 200 Msps,
             the 12 G-sample job cut into 8 time slices shared by the N ranks, with the SURVEY 8d parity
             windows), cfg5 (const i16 sweep, six rates + an irregular ratio), each with the oracle's
             1-core and all-core rate beside it
cpu_baseline = the oracle (C restatement of the reference loop + the reference's own complex.c
             when oracle/_ref exists) on the host cores, bounded sample, rank 0 at N=1 only

`--impl reference` times that CPU path alone (all host threads) and prints the same JSON shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "IQ Msamples/s through NCO mixer"
UNIT = "Msamples/s"
FS = 10_000_000
SHIFT = 100000.0
SECONDS_PER_GPU = 64
BYTES_PER_SAMPLE = 12  # f32 in (8) + i16 out (4)   
CPU_SAMPLE_CAP = 240_000_000


def workload_config(n_gpus):
    return {
        "workload": "const f32->i16 @ 10 Msps, --shift 100000 (BASELINE configs[1]), "
                    f"{SECONDS_PER_GPU} s of stream per GPU = {SECONDS_PER_GPU * FS // 10**6} Msamples/GPU/step",
        "intype": "f32", "outtype": "i16", "samplerate": FS, "shift_hz": SHIFT,
        "samples_per_gpu_per_step": SECONDS_PER_GPU * FS,
        "partition": f"{n_gpus} contiguous time slices, analytic samplenum seed, no collective",
        "l2": "inputs (5.12 GB/GPU) exceed L2 (126 MB); no flush needed",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the steps run (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1, note):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, power, reasons = [], [], [], set()
        for ts, line in list(self.lines):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax.append(mx)
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                try:
                    power.append(float(f[3]))
                except ValueError:
                    pass
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm), "window": note}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


def bind_near_gpu(local):
    """Pin this process to the CPUs NVML reports as local to its GPU (same NUMA node / PCIe root), so that the
    pinned host buffers of the e2e leg are first-touched next to the GPU.  With 8 ranks pulling ~50 GB/s each
    from host memory, buffers on the far socket halve the end-to-end rate.  Returns the previous affinity
    (to restore for the CPU baseline) or None when NVML / the cpuset does not allow it."""
    try:
        import pynvml
        import torch
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1} & prev
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_baseline(threads, seconds_target=12.0):
    """The reference's CPU path (oracle) on a bounded sample of the headline workload."""
    import numpy as np
    from tests.oracle_lib import Oracle
    oracle = Oracle()
    rng = np.random.default_rng(10_000_000)
    probe = 1_000_000
    x = rng.uniform(-0.7, 0.7, 2 * probe).astype(np.float32)
    t, _ = oracle.bench_const(x, probe, 1, 0, SHIFT, FS, threads)
    rate = probe / t
    n = int(max(probe, min(rate * seconds_target, CPU_SAMPLE_CAP)))
    x = rng.uniform(-0.7, 0.7, 2 * n).astype(np.float32)
    t, _ = oracle.bench_const(x, n, 1, 0, SHIFT, FS, threads)
    kind = "port"
    desc = (f"{n} samples of the workload (const f32->i16, 10 Msps, shift 100000) in memory, {threads} thread(s) over contiguous "
            f"time slices (the reference itself is single-threaded), C restatement of dsp.rs:117-134 + main.rs:73-87 calling "
            f"{'the reference complex.c compiled unmodified (oracle/_ref)' if oracle.using_ref else 'the restated ccexpf'}, "
            f"glibc {oracle.libc_version()}, gcc -O2 -ffp-contract=off")
    return {"value": n / t / 1e6, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc, "sample_samples": n}, n, t


def cpu_config_columns(threads):
    """CPU columns of the `configs` block (tools/workloads.py: the oracle on a bounded sample of every config)."""
    from tests.oracle_lib import Oracle
    from tools import workloads as W
    return W.cpu_columns(Oracle(), threads, budget_s=float(os.environ.get("DOPPLER_BENCH_CPU_SECONDS", "1.0")))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    vals = []
    t_tot = 0.0
    base = None
    n = 0
    steps = max(1, args.steps)
    per_step = max(1.0, min(12.0, 120.0 / (steps + args.warmup)))
    if os.environ.get("DOPPLER_BENCH_CPU_SECONDS"):   # tests/test_bench_contract.py: a short sample
        per_step = float(os.environ["DOPPLER_BENCH_CPU_SECONDS"])
    for i in range(args.warmup + steps):
        base, n, t = cpu_baseline(threads, seconds_target=per_step)
        if i >= args.warmup:
            vals.append(base["value"])
            t_tot += t
    v = statistics.mean(vals)
    base["value"] = v
    one, _, _ = cpu_baseline(1, seconds_target=min(per_step, 3.0))
    base["value_1core"] = one["value"]
    cfg = workload_config(args.gpus)
    cfg["reference_arm_sample"] = (f"each step mixes a bounded sample of the workload: {n} samples (about {per_step:.0f} s of CPU work on "
                                   f"{threads} threads), not the {SECONDS_PER_GPU * FS} of the GPU arm's step; the metric is a rate")
    configs = None
    if not args.no_configs:
        try:
            cols = cpu_config_columns(threads)
            configs = {"note": "CPU columns only: the oracle (reference CPU path) on a bounded sample of every BASELINE config", **cols}
        except Exception as e:   # noqa: BLE001
            configs = {"error": f"{type(e).__name__}: {e}"[:300]}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg, "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "configs": configs,
    }
    emit(line)
    return 0


_REAL_STDOUT = None


def emit(line):
    """The one JSON line, on the process's real stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


class Dist:
    """The few reductions the bench needs (device-side, NCCL when world > 1)."""

    def __init__(self, world, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def _red(self, v, op, dtype):
        import torch
        if self.world == 1:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], dtype=dtype, device=self.dev)
        dist.all_reduce(t, op=op)
        return t.item()

    def max(self, v):
        import torch
        import torch.distributed as dist
        return float(self._red(float(v), dist.ReduceOp.MAX, torch.float64))

    def sum_int(self, v):
        import torch
        import torch.distributed as dist
        return int(self._red(int(v), dist.ReduceOp.SUM, torch.int64))

    def all_true(self, v):
        import torch
        import torch.distributed as dist
        return bool(self._red(int(bool(v)), dist.ReduceOp.MIN, torch.int64))


# ---- the other BASELINE configs, full size (the `configs` block) -----------------------------------------------

def run_cfg3(mixer, stream, D, peak, oracle_threads):
    """cfg3: track replay, i16 @ 1.024 Msps, 600 s (614.4 M samples, ~600 shifts, long reset periods -> COLUMN segments).
    Every rank runs the whole config (it is a one-GPU config; N ranks = N replicas)."""
    import numpy as np
    import torch
    from tests.oracle_lib import Oracle
    from tools import workloads as W
    c = W.CFG3
    fs, n = c["fs"], c["secs"] * c["fs"]
    shifts = W.cfg_schedule(c)
    x = torch.empty(n * 4, dtype=torch.uint8, device=D.dev)
    y = torch.empty(n * 4, dtype=torch.uint8, device=D.dev)
    W.fill_device(x, W.I16)

    def call():
        return mixer.mix_blocks_dev(x.data_ptr(), x.numel(), W.I16, W.I16, shifts, fs, 0, y.data_ptr(), y.numel(), stream=stream.cuda_stream)

    l0 = mixer.launch_count
    for _ in range(3):
        call()
    D.barrier()
    iters, times = 5, []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        e0.record(stream)
        for _ in range(iters):
            call()
        e1.record(stream)
        D.barrier()
        times.append(D.max(e0.elapsed_time(e1) / iters))
    single = []
    for _ in range(3):   # one call alone: host planning + launch + kernel, nothing to overlap with
        t0 = time.perf_counter()
        call()
        t1 = time.perf_counter()
        stream.synchronize()
        single.append((t1 - t0, time.perf_counter() - t0))
    ms = statistics.median(times)
    # parity: the first and the last second of the stream in full, plus 22 random 65536-sample windows on block boundaries
    oracle = Oracle()
    bs = 2048
    rng = np.random.default_rng(3 + D.rank)
    wins = [(0, fs), (n - fs, n)] + [(int(s) * bs, int(s) * bs + 65536) for s in rng.integers(1, (n - 65536) // bs, 22)]
    ok, checked, osecs = W.check_windows(oracle, x, y, W.I16, W.I16, shifts, fs, 0, wins, oracle_threads)
    ok = D.all_true(ok)
    launches = mixer.launch_count - l0
    del x, y
    torch.cuda.empty_cache()
    return {"workload": "track replay i16->i16 @ 1.024 Msps, 600 s analytic overpass (f_tx 437.505 MHz, offset 5000 Hz), one call per step",
            "samples_per_gpu": n, "ms_per_call": ms, "msps": D.world * n / ms / 1e3, "gbs_per_gpu": n * 8 / ms / 1e6,
            "frac": n * 8 / ms / 1e6 / peak, "bytes_per_sample": 8, "host_plan_ms": min(a for a, _ in single) * 1e3,
            "single_call_ms": min(b for _, b in single) * 1e3, "timing": "median of 3 batches of 5 calls queued back to back, CUDA events on the launch stream, max over ranks",
            "parity_ok": ok, "parity_samples_checked": checked, "parity_windows": "first and last second in full + 22 random 65536-sample windows per rank, oracle seeded analytically, end state compared",
            "gpu_launches_per_call": launches // (3 + 15 + 3), "scaling": "replicas (a 1-GPU config)"}


def run_cfg4(mixer, stream, D, peak, oracle_threads):
    """cfg4: track f32->f32 @ 200 Msps, 60 s overpass = 12 G samples, cut into 8 contiguous time slices; the N ranks
    share the slices (rank r mixes slices [8r/N, 8(r+1)/N) one after the other through one 24 GB buffer pair), so the
    WHOLE job is done at every N (strong scaling).  Parity as SURVEY 8d specifies: 2^24 samples straddling every slice
    boundary (the last 2^23 of one slice + the first 2^23 of the next) plus the first and the last second in full."""
    import torch
    from tests.oracle_lib import Oracle
    from tools import workloads as W
    c = W.CFG4
    fs, total, ns = c["fs"], c["secs"] * c["fs"], c["slices"]
    bs = 1024
    shifts = W.cfg_schedule(c)
    begins, seeds = W.dsp.slice_seeds(0, shifts, bs, fs, total, ns)
    per = max(1, ns // D.world)
    mine = list(range(D.rank * per, min(ns, (D.rank + 1) * per))) if D.rank * per < ns else []
    cap = max(begins[i + 1] - begins[i] for i in range(ns))
    x = torch.empty(cap * 8, dtype=torch.uint8, device=D.dev)
    y = torch.empty(cap * 8, dtype=torch.uint8, device=D.dev)
    W.fill_device(x, W.F32)

    def run_slice(i):
        n = begins[i + 1] - begins[i]
        return mixer.mix_blocks_dev(x.data_ptr(), n * 8, W.F32, W.F32, shifts[begins[i] // bs:], fs, seeds[i], y.data_ptr(), n * 8,
                                    stream=stream.cuda_stream)

    l0 = mixer.launch_count
    chain_ok = True
    for i in mine:   # warm-up pass (also: the carried state equals the analytic seed of the next slice)
        chain_ok = chain_ok and run_slice(i) == seeds[i + 1]
    D.barrier()
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        e0.record(stream)
        for i in mine:
            run_slice(i)
        e1.record(stream)
        D.barrier()
        times.append(D.max(e0.elapsed_time(e1)))
    ms = statistics.median(times)
    launches = mixer.launch_count - l0
    # parity windows, slice by slice (the output buffer is reused)
    oracle = Oracle()
    half = 1 << 23
    ok, checked = chain_ok, 0
    for i in mine:
        n = begins[i + 1] - begins[i]
        run_slice(i)
        stream.synchronize()
        wins = [(0, half), ((n - half) // bs * bs, n)]
        if i == 0:
            wins[0] = (0, fs)               # the first second in full
        if i == ns - 1:
            wins[1] = ((n - fs) // bs * bs, n)   # the last second in full
        o, cnt, _ = W.check_windows(oracle, x, y, W.F32, W.F32, shifts, fs, begins[i], wins, oracle_threads)
        ok, checked = ok and o, checked + cnt
    ok = D.all_true(ok)
    checked = D.sum_int(checked)
    active = min(D.world, ns)
    del x, y
    torch.cuda.empty_cache()
    return {"workload": "track f32->f32 @ 200 Msps, 60 s analytic overpass (f_tx 4.2 GHz): 12 G samples in 8 contiguous time slices "
                        f"on 8192-byte block boundaries, analytic samplenum seeds, no collective; {D.world} rank(s) x {len(mine)} slice(s) each",
            "samples_total": total, "slices": ns, "slices_per_rank": len(mine), "ms_whole_job": ms, "msps": total / ms / 1e3,
            "gbs_per_gpu": total * 16 / active / ms / 1e6, "frac": total * 16 / active / ms / 1e6 / peak, "bytes_per_sample": 16,
            "timing": "the rank's slices queued back to back, CUDA events on the launch stream, median of 3 passes, max over ranks",
            "parity_ok": ok, "parity_samples_checked": checked,
            "parity_windows": "per slice: first 2^23 and last 2^23 samples (2^24 straddling every slice boundary); first and last second of the job in full; "
                              "bit-exact f32 vs the oracle, carried samplenum == analytic seed of the next slice",
            "gpu_launches_per_pass": launches // 4, "scaling": "strong (the whole job at every N)"}


def run_cfg5(mixer, stream, D, peak):
    """cfg5: const i16->i16, 1 s buffers at six rates (r = -15/256, P = 256) + one irregular ratio; every launch timed on
    its own with CUDA events, L2 flushed before launches whose buffers could sit in it; every rank runs the same (weak)."""
    import torch
    from tools import workloads as W
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=D.dev)
    points = [(str(fs), fs, W.cfg5_shift(fs)) for fs in W.CFG5_RATES] + [("irregular", W.CFG5_IRREGULAR["fs"], W.CFG5_IRREGULAR["shift"])]
    out = {}
    for key, fs, shift in points:
        n = fs
        x = torch.empty(n * 4, dtype=torch.uint8, device=D.dev)
        y = torch.empty(n * 4, dtype=torch.uint8, device=D.dev)
        W.fill_device(x, W.I16)
        need_flush = 2 * x.numel() < 2 * 126e6
        iters = 20
        times = []
        with torch.cuda.stream(stream):
            for i in range(3 + iters):
                if need_flush:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                mixer.mix_dev(x.data_ptr(), x.numel(), W.I16, W.I16, shift, fs, 0, y.data_ptr(), y.numel(), stream=stream.cuda_stream)
                e1.record(stream)
                stream.synchronize()
                if i >= 3:
                    times.append(e0.elapsed_time(e1) * 1e-3)
        med, best = D.max(statistics.median(times)), D.max(min(times))
        out[key] = {"samplerate": fs, "shift_hz": shift, "samples_per_gpu": n, "median_us": med * 1e6, "best_us": best * 1e6,
                    "msps": D.world * n / med / 1e6, "frac": n * 8 / med / 1e9 / peak, "l2_flushed": need_flush}
        del x, y
        torch.cuda.empty_cache()
    out["note"] = ("1 s of stream per launch, 20 launches each timed alone with CUDA events after 3 warm-ups (median; max over ranks); "
                   "frac = 8 B/sample / time / measured HBM peak; buffers below 2x L2 are preceded by a 512 MB memset")
    return out


def run_f4(mixer, stream, D, peak):
    """SURVEY 8(f) row 4, the fused downstream stage: const f32 -> i16 @ 10 Msps (the headline's input) followed by a
    decimate-by-8 49-tap low-pass, one pass (doppler_b200_mix_decimate).  Device-resident rate, the host-buffer rate (D2H
    shrinks by 8) and a bit-exact check of a window against the oracle's specification."""
    import ctypes

    import numpy as np
    import torch

    import doppler_b200
    from doppler_b200 import F32, I16, _lib
    from tests.oracle_lib import Oracle
    M, ntaps = 8, 49
    t = np.arange(ntaps) - (ntaps - 1) / 2.0
    h = np.sinc(2 * 0.05 * t) * np.hamming(ntaps)
    taps = (h / h.sum()).astype(np.float32)
    dec = doppler_b200.Decimator(mixer, taps, M)
    n = 256_000_000
    x = torch.empty(2 * n, dtype=torch.float32, device=D.dev)
    x.uniform_(-0.7, 0.7)
    y = torch.empty(2 * (n // M + 2), dtype=torch.int16, device=D.dev)
    torch.cuda.synchronize()
    times = []
    for i in range(2 + 5):
        dec.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        nbytes, sn = dec.mix_dev(x.data_ptr(), 8 * n, F32, I16, SHIFT, FS, 0, y.data_ptr(), 2 * y.numel(), stream=stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1) * 1e-3)
    med = D.max(statistics.median(times))
    # parity: the first 2^20 inputs against the specification
    w = 1 << 20
    want, st = Oracle().mix_decimate(x[:2 * w].cpu().numpy().view(np.uint8), F32, I16, SHIFT, FS, taps, M)
    ok = bool(np.array_equal(y[:want.size // 2].cpu().numpy().view(np.uint8), want))
    del x, y
    torch.cuda.empty_cache()
    # host buffers (pinned): H2D 8 B/sample, D2H 4/8 B/sample
    lib = _lib.load()
    ne = 160_000_000
    hin, hout = lib.doppler_b200_host_alloc(8 * ne), lib.doppler_b200_host_alloc(4 * (ne // M + 2))
    a_in = np.ctypeslib.as_array(ctypes.cast(hin, ctypes.POINTER(ctypes.c_float)), shape=(2 * ne,))
    a_in[:] = np.random.default_rng(4 + D.rank).uniform(-0.7, 0.7, 2 * ne).astype(np.float32)
    tt = []
    for i in range(1 + 3):
        dec.reset()
        snc, got = ctypes.c_uint32(0), ctypes.c_size_t(0)
        D.barrier()
        t0 = time.perf_counter()
        rc = lib.doppler_b200_mix_decimate(dec._d, hin, 8 * ne, F32, I16, ctypes.c_float(SHIFT), FS, ctypes.byref(snc), hout, 4 * (ne // M + 2), ctypes.byref(got))
        dt = D.max(time.perf_counter() - t0)
        if rc != 0:
            raise RuntimeError(f"doppler_b200_mix_decimate failed rc={rc}")
        if i >= 1:
            tt.append(dt)
    lib.doppler_b200_host_free(hin)
    lib.doppler_b200_host_free(hout)
    dec.close()
    bps = 8 + 4.0 / M
    return {"workload": f"const f32 @ 10 Msps --shift 100000 -> mixer -> {ntaps}-tap low-pass, decimate by {M} -> i16 (one pass)",
            "samples_per_gpu": n, "ms_per_call": med * 1e3, "msps_in": D.world * n / med / 1e6, "bytes_per_input_sample": bps,
            "frac": n * bps / med / 1e9 / peak, "parity_ok": D.all_true(ok), "parity_samples_checked": w,
            "e2e_msps_in": D.world * ne / statistics.median(tt) / 1e6, "e2e_d2h_bytes": 4 * (ne // M),
            "kernel": "dmix::mix_decimate_fast_kernel (4 consecutive outputs per thread, taps as uniform-register FFMA2 operands from the kernel parameters)",
            "note": "not in the reference (SURVEY 8f row 4); specification = oracle_mix_decimate; bound by issue slots + the shared-memory pipe, not HBM (profiles/r02_ncu_decim_f32_i16.txt)"}


def run_per_block():
    """The reference's own call granularity: one 8192-byte block (2048 i16 samples) per host call (main.rs:49,70), pageable
    caller buffers, measured by the C harness tools/tune/percall (no Python in the loop): resident kernel (default) against one
    zero-copy launch per block."""
    exe = os.path.join(ROOT, "tools", "tune", "percall")
    if not os.path.exists(exe):
        return {"error": "tools/tune/percall is not built (python -c 'import __graft_entry__ as g; g.build()')"}
    r = subprocess.run([exe], env=dict(os.environ, PERCALL_QUICK="1"), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=120)
    rows = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    out = {"workload": "doppler_b200_mix i16->i16, 2048 samples (8192 bytes) per call, shift 5000 Hz @ 1.024 Msps, pageable buffers, 5000 calls",
           "reference_cpu_us_per_block": "~80 (25 Msample/s on one core, cpu_baseline.value_1core)"}
    for row in rows:
        key = "resident_kernel" if row["path"].startswith("resident") else "launch_per_block"
        if "paced" in row:   # calls 30-60 us apart (a stream, not a file): the call's own duration
            out.setdefault("paced_" + key, {"pacing": row["paced"], "mean_us": row["us_per_call"], "median_us": row["median_us"], "p99_us": row["p99_us"]})
        else:
            out[key] = {"us_per_call": row["us_per_call"], "msps": row["msps"]}
    return out


def run_cfg1_cli(oracle_threads):
    """cfg1: `doppler const -s 256000 -i i16 --shift -15000` over 1 s of i16 IQ through the CLI (stdin -> stdout), bytes
    compared with the oracle's restatement of the reference's const driver.  Wall clock of the whole process (CUDA
    start-up included) beside the oracle's in-memory time for the same stream."""
    import numpy as np
    from tests.oracle_lib import Oracle
    cli = os.path.join(ROOT, "doppler_b200", "bin", "doppler")
    n, fs = 256_000, 256_000
    rng = np.random.default_rng(20150122)
    t = np.arange(n)
    sig = 0.25 * np.exp(2j * np.pi * 15000.0 / fs * t) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty(2 * n, dtype="<i2")
    iq[0::2] = np.clip(np.round(sig.real * 32767), -32768, 32767)
    iq[1::2] = np.clip(np.round(sig.imag * 32767), -32768, 32767)
    raw = iq.view(np.uint8)
    oracle = Oracle()
    t0 = time.perf_counter()
    want, _, panicked = oracle.const_stream(raw, 0, 0, -15000, fs)
    cpu_s = time.perf_counter() - t0
    walls = []
    ok = not panicked
    for _ in range(2):
        t0 = time.perf_counter()
        r = subprocess.run([cli, "const", "-s", str(fs), "-i", "i16", "--shift", "-15000"], input=raw.tobytes(), capture_output=True, timeout=120)
        walls.append(time.perf_counter() - t0)
        ok = ok and r.returncode == 0 and r.stdout == want.tobytes()
    return {"workload": "CLI: doppler const -s 256000 -i i16 --shift -15000, 256000 samples on stdin (BASELINE configs[0])",
            "cli_wall_s": min(walls), "cli_msps": n / min(walls) / 1e6, "cpu_oracle_s": cpu_s, "cpu_oracle_msps": n / cpu_s / 1e6,
            "parity_ok": bool(ok), "note": "process wall clock incl. CUDA context creation; the oracle figure is the in-memory const driver on 1 core"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=int, default=SECONDS_PER_GPU, help="seconds of 10 Msps stream per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1/cfg3/cfg4/cfg5 block")
    ap.add_argument("--no-ncu", action="store_true", help="do not spawn ncu for roofline.traffic (use the committed capture)")
    ap.add_argument("--sustained-seconds", type=float, default=2.5)
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything libraries print there meanwhile (NCCL's version banner,
    # torchrun children ...) is sent to stderr by pointing fd 1 at fd 2 until the line is ready
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(3, args.warmup)

    import numpy as np
    import torch
    import torch.distributed as dist

    import doppler_b200
    from doppler_b200 import F32, I16

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    D = Dist(world, dev)
    D.rank = rank
    mixer = doppler_b200.Mixer(local)
    oracle_threads = max(1, host_threads() // world)

    n = args.seconds * FS  # samples per GPU per step
    gen = torch.Generator(device=dev)
    gen.manual_seed(10_000_000 + rank)
    x = torch.empty(2 * n, dtype=torch.float32, device=dev)
    x.uniform_(-0.7, 0.7, generator=gen)
    y = torch.empty(2 * n, dtype=torch.int16, device=dev)
    # this rank's time slice starts at stream sample rank*n: the library carries the reference's samplenum there
    _, seeds = doppler_b200.dsp.slice_seeds(0, SHIFT, 1024, FS, world * n, world)
    seed = seeds[rank]
    # a side stream: handle 0 (torch's legacy default stream) means "the context's own stream" to
    # the C ABI, and CUDA events must be recorded on the stream the kernels are launched on
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        return mixer.mix_dev(x.data_ptr(), 8 * n, F32, I16, SHIFT, FS, seed, y.data_ptr(), 4 * n, stream=stream)

    barrier = D.barrier
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = mixer.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ev0.record(tstream)
    for _ in range(args.steps):
        step()
    ev1.record(tstream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = mixer.launch_count - l0
    ms = D.max(ms)
    launches = D.sum_int(launches)

    # ---- sustained leg: the same launches back to back for >= 2 s (power-capped regime), timed the same way ----
    est = ms / args.steps
    ks = max(args.steps, int(args.sustained_seconds * 1e3 / est) + 1)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ts0 = time.time()
    es0.record(tstream)
    for _ in range(ks):
        step()
    es1.record(tstream)
    barrier()
    ts1 = time.time()
    ms_sus = D.max(es0.elapsed_time(es1)) / ks
    clocks = sus_clocks = None
    if rank == 0:
        clocks = sampler.window(t0, ts1, "timed region + the sustained leg that follows it at once (the timed region alone is shorter than the sampling period)")
        sus_clocks = sampler.window(ts0 + 0.3, ts1, "sustained leg, first 0.3 s excluded")
        sampler.stop()

    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = measured_peak_gbs()
    # one mixer launch per step on this rank (the phasor table is built once, in warm-up)
    achieved = BYTES_PER_SAMPLE * n / (ms_per_step * 1e-3) / 1e9
    sus_achieved = BYTES_PER_SAMPLE * n / (ms_sus * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": "dmix::mix_grid_kernel<F32,I16,16,2,3> (per-warp cp.async.bulk pipelines, phasor table in shared memory)",
                "bytes_per_sample": BYTES_PER_SAMPLE,
                "note": "per-GPU; time = CUDA events on the launch stream over the timed region / launches",
                "sustained": {"achieved": sus_achieved, "frac": sus_achieved / peak, "launches": ks, "seconds": ms_sus * ks * 1e-3,
                              "msps": world * n / (ms_sus * 1e-3) / 1e6, "clocks": sus_clocks,
                              "note": "the same launch repeated back to back (no host sync) for the stated time; max over ranks; against the same burst copy peak"}}
    del x, y
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C-ABI call (pinned host memory) ----------------
    e2e = None
    if not args.no_e2e:
        import ctypes
        prev_affinity = bind_near_gpu(local)
        from doppler_b200 import _lib
        lib = _lib.load()
        ne = min(n, 16 * FS)  # 160 Msamples per step: 1.28 GB in, 0.64 GB out through PCIe
        hin = lib.doppler_b200_host_alloc(8 * ne)
        hout = lib.doppler_b200_host_alloc(4 * ne)
        if not hin or not hout:
            raise RuntimeError("pinned host allocation failed")
        a_in = np.ctypeslib.as_array(ctypes.cast(hin, ctypes.POINTER(ctypes.c_float)), shape=(2 * ne,))
        a_out = np.ctypeslib.as_array(ctypes.cast(hout, ctypes.POINTER(ctypes.c_uint8)), shape=(4 * ne,))
        a_in[:] = np.random.default_rng(10_000_000 + rank).uniform(-0.7, 0.7, 2 * ne).astype(np.float32)

        def e2e_step(pin=hin, pout=hout, view=a_out):
            sn = ctypes.c_uint32(seed)
            got = ctypes.c_size_t(0)
            rc = lib.doppler_b200_mix(mixer._ctx, pin, 8 * ne, F32, I16, ctypes.c_float(SHIFT), FS, ctypes.byref(sn), pout, 4 * ne,
                                      ctypes.byref(got))
            if rc != 0 or got.value != 4 * ne:
                raise RuntimeError(f"doppler_b200_mix failed rc={rc}")
            return int(view[0]) + int(view[-1])  # the result is read on the host

        def probe_step():
            rc = lib.doppler_b200_pipeline_probe(mixer._ctx, hin, 8 * ne, F32, I16, hout, 4 * ne)
            if rc != 0:
                raise RuntimeError(f"doppler_b200_pipeline_probe failed rc={rc}")
            return int(a_out[0]) + int(a_out[-1])

        def timed(fn, k):
            for _ in range(2):
                fn()
            barrier()
            t_0 = time.perf_counter()
            for _ in range(k):
                fn()
            torch.cuda.synchronize()
            return D.max(time.perf_counter() - t_0)

        ke = max(3, min(args.steps, 10))
        te = timed(e2e_step, ke)
        tc = timed(probe_step, ke)
        e2e = {"value": world * ne * ke / te / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * ne, "d2h_bytes_per_step": 4 * ne,
               "samples_per_gpu_per_step": ne, "steps": ke,
               "api": "doppler_b200_mix (host buffers, pinned via doppler_b200_host_alloc; 32 MiB chunks, 3-slot H2D/kernel/D2H pipeline)",
               "cpu_affinity": "GPU-local CPUs (NVML)" if prev_affinity is not None else "unchanged",
               "ceiling": {"value": world * ne * ke / tc / 1e6, "unit": UNIT, "frac_of_ceiling": tc / te,
                           "how": "doppler_b200_pipeline_probe: the same pipeline (chunks, slots, streams, buffers) with the kernel skipped -- "
                                  "what the box's host memory / PCIe path delivers to all ranks at once"}}
        # the same call on ordinary pageable memory (what a caller's Vec<u8> is): staged through pinned slots by the library
        p_in = np.random.default_rng(5 + rank).uniform(-0.7, 0.7, 2 * ne).astype(np.float32)
        p_out = np.empty(4 * ne, dtype=np.uint8)
        tp = timed(lambda: e2e_step(p_in.ctypes.data, p_out.ctypes.data, p_out), 3)
        e2e["pageable"] = {"value": world * ne * 3 / tp / 1e6, "unit": UNIT, "note": "same call, caller buffers not pinned (numpy arrays)"}
        del p_in, p_out
        lib.doppler_b200_host_free(hin)
        lib.doppler_b200_host_free(hout)
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)

    # ---- every other BASELINE config at full size ------------------------------------------------
    configs = None
    if not args.no_configs:
        configs = {"cfg2": "the headline of this line (value / roofline / e2e)"}
        t_c = time.time()

        def guarded(name, fn, *a):
            """A config that fails reports its error in place; it never costs the line its headline.  (Every rank runs the same
            code on the same inputs, so a failure is collective and the ranks stay in step.)"""
            try:
                configs[name] = fn(*a)
            except Exception as e:   # noqa: BLE001
                configs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.synchronize()

        guarded("cfg3", run_cfg3, mixer, tstream, D, peak, oracle_threads)
        guarded("cfg4", run_cfg4, mixer, tstream, D, peak, oracle_threads)
        guarded("cfg5", run_cfg5, mixer, tstream, D, peak)
        guarded("f4_mix_decimate", run_f4, mixer, tstream, D, peak)
        if rank == 0:
            guarded("cfg1", run_cfg1_cli, oracle_threads)
        configs["gpu_seconds"] = time.time() - t_c

    mixer.close()
    if configs is not None and rank == 0 and world == 1:   # (its own process and context: after this one's is gone)
        try:
            configs["per_block_host_call"] = run_per_block()
        except Exception as e:   # noqa: BLE001
            configs["per_block_host_call"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    base = None
    if rank == 0:
        threads = host_threads()
        if configs is not None:   # CPU columns beside every config (rank 0's host cores; the other ranks are idle here)
            t_c = time.time()
            try:
                cols = cpu_config_columns(threads)
                for k in ("cfg1", "cfg3", "cfg4"):
                    if isinstance(configs.get(k), dict):
                        configs[k].update(cols[k])
                for k, v in cols["cfg5"].items():
                    if isinstance(configs.get("cfg5"), dict) and isinstance(configs["cfg5"].get(k), dict):
                        configs["cfg5"][k].update(v)
            except Exception as e:   # noqa: BLE001
                configs["cpu_columns_error"] = f"{type(e).__name__}: {e}"[:300]
            configs["cpu_seconds"] = time.time() - t_c
            configs["cpu_note"] = ("cpu_msps_*: the oracle (reference CPU path; the reference is single-threaded, the all-core figure runs "
                                   "contiguous time slices on every host thread) on a bounded in-memory sample of the same config")
        if world == 1 and not args.no_cpu_baseline:
            try:
                base, _, _ = cpu_baseline(threads)
                one, _, _ = cpu_baseline(1, seconds_target=4.0)
                base["value_1core"] = one["value"]
            except Exception as e:   # noqa: BLE001
                base = {"error": f"{type(e).__name__}: {e}"[:300]}
        # DRAM traffic of one launch at THIS launch size, by ncu, after every timed region (one GPU only: ncu serialises)
        tr = None
        if world == 1 and not args.no_ncu:
            try:
                from tools import ncu_traffic
                tr = ncu_traffic.measure(n, timeout=240, keep_csv=os.path.join(ROOT, "gpurun_out", "bench_traffic_ncu.csv"))
                roofline["traffic"] = tr["bytes"]
                roofline["traffic_note"] = (f"ncu dram__bytes_read.sum + dram__bytes_write.sum of one {n}-sample launch, captured by this run after timing "
                                            f"({tr['bytes'] / n:.3f} B/sample for {BYTES_PER_SAMPLE} algorithmic; read {tr['bytes_read']:.4g}, write {tr['bytes_write']:.4g})")
            except Exception as e:   # no ncu / no permission on this box: fall back to the committed capture, and say so
                roofline["traffic_error"] = str(e)[:200]
        if roofline["traffic"] is None:
            try:
                com = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                per_sample = com.get("mix_f32_i16_bytes_per_sample")
                if per_sample:
                    roofline["traffic"] = per_sample * n
                    roofline["traffic_note"] = (f"committed capture, not this run: {per_sample} B/sample (profiles/traffic.json, "
                                                f"{com.get('mix_f32_i16_samples', '?')}-sample launch) x {n} samples")
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world), "roofline": roofline, "cpu_baseline": base, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks, "configs": configs,
            "libm_compatible": bool(doppler_b200.dsp.libm_compatible()),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
