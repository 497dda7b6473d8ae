#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the NCO mixer on N B200s (BASELINE.json's metric).

A "step" is one pass of the hot path over one batch of synthetic IQ.  Workload (configs[1] of
BASELINE.json): const mode, f32 -> i16, fs = 10 Msps, --shift 100000; the batch is 64 s of that
stream per GPU (640 M complex samples: 5.12 GB in, 2.56 GB out -- far larger than the 126 MB L2,
so no flush is needed between iterations).  With N > 1 (torchrun, one rank per GPU) the stream is
N x 64 s long, cut into contiguous time slices; every rank seeds its slice with the analytically
carried samplenum (doppler_b200_samplenum_advance) -- no collective on the data path ("weak").

value      = samples all ranks processed / max-over-ranks device time, inputs resident in HBM
e2e        = same metric through the host-buffer C-ABI call (pinned host in/out, H2D + kernel +
             D2H inside the timed region)
roofline   = algorithmic bytes (12 B/sample for f32->i16) / CUDA-event time of the mixer launch,
             against MEASURED_PEAKS.json hbm_gbs
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
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

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

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
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
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: use everything we saw
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "timed region + 0.7 s of the same steps (the timed region alone is shorter than the sampling period)"}


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


def cpu_baseline(threads, seconds_target=12.0):
    """The reference's CPU path (oracle) on a bounded sample of the same workload."""
    import numpy as np
    from tests.oracle_lib import Oracle
    oracle = Oracle()
    rng = np.random.default_rng(10_000_000)
    probe = 1_000_000
    x = rng.uniform(-0.7, 0.7, 2 * probe).astype(np.float32)
    t, _ = oracle.bench_const(x, probe, 1, 0, SHIFT, FS, threads)
    rate = probe / t
    n = int(max(probe, min(rate * seconds_target, 240_000_000)))
    x = rng.uniform(-0.7, 0.7, 2 * n).astype(np.float32)
    t, _ = oracle.bench_const(x, n, 1, 0, SHIFT, FS, threads)
    kind = "port"
    desc = (f"{n} samples of the workload (const f32->i16, 10 Msps, shift 100000) in memory, {threads} thread(s), "
            f"C restatement of dsp.rs:117-134 + main.rs:73-87 calling "
            f"{'the reference complex.c compiled unmodified (oracle/_ref)' if oracle.using_ref else 'the restated ccexpf'}, "
            f"glibc {oracle.libc_version()}, gcc -O2 -ffp-contract=off")
    return {"value": n / t / 1e6, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc}, n, t


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    vals = []
    t_tot = 0.0
    base = None
    steps = max(1, args.steps)
    per_step = max(1.0, min(12.0, 150.0 / (steps + args.warmup)))
    if os.environ.get("DOPPLER_BENCH_CPU_SECONDS"):   # tests/test_bench_contract.py: a short sample
        per_step = float(os.environ["DOPPLER_BENCH_CPU_SECONDS"])
    for i in range(args.warmup + steps):
        base, n, t = cpu_baseline(threads, seconds_target=per_step)
        if i >= args.warmup:
            vals.append(base["value"])
            t_tot += t
    v = statistics.mean(vals)
    base["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus), "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=int, default=SECONDS_PER_GPU, help="seconds of 10 Msps stream per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    mixer = doppler_b200.Mixer(local)

    n = args.seconds * FS  # samples per GPU per step
    gen = torch.Generator(device=dev)
    gen.manual_seed(10_000_000 + rank)
    x = torch.empty(2 * n, dtype=torch.float32, device=dev)
    x.uniform_(-0.7, 0.7, generator=gen)
    y = torch.empty(2 * n, dtype=torch.int16, device=dev)
    # this rank's time slice starts at stream sample rank*n: carry the reference's samplenum there
    seed = doppler_b200.samplenum_advance(0, SHIFT, FS, rank * n)
    # a side stream: handle 0 (torch's legacy default stream) means "the context's own stream" to
    # the C ABI, and CUDA events must be recorded on the stream the kernels are launched on
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        return mixer.mix_dev(x.data_ptr(), 8 * n, F32, I16, SHIFT, FS, seed, y.data_ptr(), 4 * n, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    # The timed region lasts ~K ms, shorter than nvidia-smi's sampling period: keep the SAME steps running
    # (untimed) for a moment so that the clock / throttle record is taken under this workload's load.
    if rank == 0:
        t_hold = time.time() + 0.7
        while time.time() < t_hold:
            for _ in range(20):
                step()
            torch.cuda.synchronize()
    t1 = time.time()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = measured_peak_gbs()
    # one mixer launch per step on this rank (the phasor table is built once, in warm-up)
    achieved = BYTES_PER_SAMPLE * n / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": "dmix::mix_grid_kernel<F32,I16,16,2,3> (per-warp cp.async.bulk pipelines, phasor table in shared memory)", "bytes_per_sample": BYTES_PER_SAMPLE,
                "note": "per-GPU; time = CUDA events on the launch stream over the timed region / launches"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            per_sample = json.load(open(tr)).get("mix_f32_i16_bytes_per_sample")
            if per_sample:
                roofline["traffic"] = per_sample * n  # bytes per launch, from the committed ncu --set full capture
                roofline["traffic_note"] = f"{per_sample} B/sample (profiles/traffic.json) x {n} samples per launch"
        except Exception:
            pass

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

        def e2e_step():
            sn = ctypes.c_uint32(seed)
            got = ctypes.c_size_t(0)
            rc = lib.doppler_b200_mix(mixer._ctx, hin, 8 * ne, F32, I16, ctypes.c_float(SHIFT), FS, ctypes.byref(sn), hout, 4 * ne,
                                      ctypes.byref(got))
            if rc != 0 or got.value != 4 * ne:
                raise RuntimeError(f"doppler_b200_mix failed rc={rc}")
            return int(a_out[0]) + int(a_out[-1])  # the result is read on the host

        for _ in range(2):
            e2e_step()
        barrier()
        te0 = time.perf_counter()
        ke = max(3, min(args.steps, 10))
        for _ in range(ke):
            e2e_step()
        torch.cuda.synchronize()
        te = time.perf_counter() - te0
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": world * ne * ke / te / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * ne, "d2h_bytes_per_step": 4 * ne,
               "samples_per_gpu_per_step": ne, "steps": ke,
               "api": "doppler_b200_mix (host buffers, pinned via doppler_b200_host_alloc; 32 MiB chunks, 3-slot H2D/kernel/D2H pipeline)",
               "cpu_affinity": "GPU-local CPUs (NVML)" if prev_affinity is not None else "unchanged"}
        lib.doppler_b200_host_free(hin)
        lib.doppler_b200_host_free(hout)
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        base, _, _ = cpu_baseline(threads)
        one, _, _ = cpu_baseline(1, seconds_target=4.0)
        base["value_1core"] = one["value"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world), "roofline": roofline, "cpu_baseline": base, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
        }
        emit(line)
    mixer.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
