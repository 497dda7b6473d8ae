#!/usr/bin/env python
"""tools/fuzz_parity.py -- randomized parity fuzz of the planned entry point against the oracle (GPU box).

Random per-block schedules (shifts with short, long and no reset periods, |r| > 1, tiny r), random lengths
from a few samples to ~12 M (so that GRID, COLUMN and slow tiles, the 4 Mi-sample COLUMN threshold and
several host-pipeline chunks all occur), random start samplenum, all four type pairs.  Each trial runs on one of
five code paths: the product's thresholds (small kernel / zero-copy path for short inputs), the bulk-async kernels
only, a two-context device group (time slices with analytic seeds), the fused mix + decimating FIR, and one call per
8192-byte block (the resident kernel).
Exits non-zero at the first mismatch and prints the reproducer."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doppler_b200  # noqa: E402
from doppler_b200 import F32, I16  # noqa: E402
from tests.oracle_lib import BUFFER_SIZE, Oracle, same_bits_f32  # noqa: E402

BPS = {I16: 4, F32: 8}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=120)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    oracle, mixer = Oracle(), doppler_b200.Mixer(0)
    bulk = doppler_b200.Mixer(0)
    bulk.tune(small_max_samples=0, tiny_host_bytes=0)
    group = doppler_b200.MultiMixer([0, 0])
    rng = np.random.default_rng(args.seed)
    pool = np.array([-9876.54, 7321.7, 5000.0, -3211.11, -15000.0, 0.0, 12_345.678, 1.0, 815000.0, -1234.5, 48000.0, 0.37,
                     -250_000.0, 3.0e6, 1e-3], dtype=np.float32)
    total = 0
    for t in range(args.trials):
        intype, outtype = [(I16, I16), (I16, F32), (F32, I16), (F32, F32)][t % 4]
        fs = int(rng.choice([8000, 96_000, 1_024_000, 2_400_000, 200_000_000]))
        nruns = int(rng.integers(1, 6))
        big = rng.random() < 0.35
        shifts = np.concatenate([np.repeat(rng.choice(pool) + np.float32(rng.normal(0, 3)), int(rng.integers(1, 1500 if big else 60)))
                                 for _ in range(nruns)]).astype(np.float32)
        nbytes = shifts.size * BUFFER_SIZE - BPS[intype] * int(rng.integers(0, BUFFER_SIZE // BPS[intype]))
        start = int(rng.choice([0, 1, 5, 77_777, 2**24 + 5, 2**31 - 9, 2**32 - 3]))
        n = nbytes // BPS[intype]
        if intype == I16:
            buf = rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype(np.int16).view(np.uint8)
        else:
            buf = rng.uniform(-1.2, 1.2, 2 * n).astype(np.float32).view(np.uint8)
        path = ("default", "bulk", "group", "decimate", "perblock")[(t // 4) % 5]
        if path == "perblock":                                   # the reference's call pattern: one call per 8192-byte block
            shifts = shifts[:200]
            buf = buf[:shifts.size * BUFFER_SIZE]
            n = buf.size // BPS[intype]
        if path == "decimate":
            M, ntaps = int(rng.integers(1, 12)), int(rng.integers(1, 70))
            taps = rng.uniform(-0.3, 0.3, ntaps).astype(np.float32)
            dec = doppler_b200.Decimator(mixer, taps, M)
            cut = (n // 2) // (BUFFER_SIZE // BPS[intype]) * (BUFFER_SIZE // BPS[intype])      # two calls, cut on a block boundary
            bpc = cut * BPS[intype]
            g1, sn = dec.mix_blocks(buf[:bpc], intype, outtype, shifts, fs, samplenum=start)
            g2, sn = dec.mix_blocks(buf[bpc:], intype, outtype, shifts[cut // (BUFFER_SIZE // BPS[intype]):], fs, samplenum=sn)
            got = np.concatenate([g1, g2])
            dec.close()
            want, st = oracle.mix_decimate(buf, intype, outtype, shifts, fs, taps, M, {"samplenum": start, "hist": np.zeros(2 * max(ntaps - 1, 1), dtype=np.float32), "pos": 0})
            sn_ref = st["samplenum"]
        elif path == "perblock":
            parts, sn = [], start
            for b in range(0, buf.size, BUFFER_SIZE):            # (served by the resident kernel after the first block)
                g, sn = mixer.mix(buf[b:b + BUFFER_SIZE], intype, outtype, float(shifts[b // BUFFER_SIZE]), fs, samplenum=sn)
                parts.append(g)
            got = np.concatenate(parts)
            want, sn_ref = oracle.mix_blocks_threads(buf, intype, outtype, shifts, fs, samplenum=start)
        else:
            m = {"default": mixer, "bulk": bulk, "group": group}[path]
            got, sn = m.mix_blocks(buf, intype, outtype, shifts, fs, samplenum=start)
            want, sn_ref = oracle.mix_blocks_threads(buf, intype, outtype, shifts, fs, samplenum=start)
        ok = sn == sn_ref and got.size == want.size and (np.array_equal(got, want) if outtype == I16 else same_bits_f32(got, want))
        total += n
        if not ok:
            bad = np.flatnonzero(got[:min(got.size, want.size)] != want[:min(got.size, want.size)])
            print(f"MISMATCH trial {t} path {path} seed {args.seed}: types {intype}->{outtype} fs {fs} start {start} n {n} shifts {np.unique(shifts)[:8]} "
                  f"sn {sn} vs {sn_ref}, first differing byte {bad[:3]}")
            return 1
    print(f"fuzz ok: {args.trials} trials, {total} samples, seed {args.seed}")
    mixer.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
