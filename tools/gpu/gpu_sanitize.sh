#!/bin/bash
# compute-sanitizer over the kernels' paths (small inputs): memcheck, racecheck (shared-memory hazards), synccheck
OUT=gpurun_out/${1:-r02san}
mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import doppler_b200
from doppler_b200 import F32, I16
from tests.oracle_lib import Oracle
o = Oracle()
m = doppler_b200.Mixer(0)
rng = np.random.default_rng(5)
bad = 0
for it, ot in [(I16, I16), (I16, F32), (F32, I16), (F32, F32)]:
    cases = [(-15000.0, 256000, 70_001), (-3_912_345.25, 200_000_000, 200_003), (7321.7, 1_024_000, 130_001),
             (1.0, 2_000_000_000, 50_001)]
    if it == ot:
        cases.append((-9876.54, 1_024_000, 4_300_003))   # >= 4 Mi samples, long period: COLUMN segments, dynamic unit claiming
    for shift, fs, n in cases:
        if it == I16:
            buf = rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype(np.int16).view(np.uint8)
        else:
            buf = rng.uniform(-0.7, 0.7, 2 * n).astype(np.float32).view(np.uint8)
        got, sn = m.mix(buf, it, ot, shift, fs)
        want, sn_ref = o.mix(buf, it, ot, shift, fs)
        ok = sn == sn_ref and np.array_equal(got, want)
        bad += not ok
        print(it, ot, shift, n, "ok" if ok else "MISMATCH")
# round 2: the bulk-async kernels on the same short inputs (small kernel and zero-copy path off), the plateau path of direct
# evaluation (samplenum above 2^24), the zero-copy per-block host path, the multi-context group, the fused decimator
m.tune(small_max_samples=0, tiny_host_bytes=0)
for it, ot in [(I16, I16), (F32, I16)]:
    for shift, fs, n, sn0 in [(-15000.0, 256000, 70_001, 0), (1.0, 2_000_000_000, 300_001, 2**26 + 5), (7321.7, 1_024_000, 130_001, 0)]:
        buf = (rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype(np.int16) if it == I16 else rng.uniform(-0.7, 0.7, 2 * n).astype(np.float32)).view(np.uint8)
        got, sn = m.mix(buf, it, ot, shift, fs, samplenum=sn0)
        want, sn_ref = o.mix(buf, it, ot, shift, fs, samplenum=sn0)
        ok = sn == sn_ref and np.array_equal(got, want)
        bad += not ok
        print("bulk", it, ot, shift, n, "ok" if ok else "MISMATCH")
m.tune(small_max_samples=4 << 20, tiny_host_bytes=128 << 10)
for n in (1, 2048, 2049, 30_001):
    buf = rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype(np.int16).view(np.uint8)
    got, sn = m.mix(buf, I16, I16, 5000.0, 1_024_000)
    want, sn_ref = o.mix(buf, I16, I16, 5000.0, 1_024_000)
    ok = sn == sn_ref and np.array_equal(got, want)
    bad += not ok
    print("tiny", n, "ok" if ok else "MISMATCH")
g = doppler_b200.MultiMixer([0, 0])
buf = rng.uniform(-0.7, 0.7, 2 * 200_003).astype(np.float32).view(np.uint8)
got, sn = g.mix(buf, F32, I16, 100000.0, 10_000_000)
want, sn_ref = o.mix(buf, F32, I16, 100000.0, 10_000_000)
ok = sn == sn_ref and np.array_equal(got, want)
bad += not ok
print("multi", "ok" if ok else "MISMATCH")
g.close()
taps = (np.hamming(33) / np.hamming(33).sum()).astype(np.float32)
d = doppler_b200.Decimator(m, taps, 8)
st, sn = None, 0
for n in (5, 70_001, 200_003):
    buf = rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype(np.int16).view(np.uint8)
    got, sn = d.mix(buf, I16, I16, 7321.7, 1_024_000, samplenum=sn)
    want, st = o.mix_decimate(buf, I16, I16, 7321.7, 1_024_000, taps, 8, st)
    ok = sn == st["samplenum"] and np.array_equal(got, want)
    bad += not ok
    print("decimate", n, "ok" if ok else "MISMATCH")
d.close()
# the register-blocked decimator (tabled shifts), both CTA sizes, and the generic kernel on the same filter
for it, ot, shift, fs, M, ntaps in [(F32, I16, 100000.0, 10_000_000, 8, 49), (I16, F32, -15000.0, 256000, 4, 33), (F32, F32, -15000.0, 256000, 5, 12)]:
    t = np.arange(ntaps) - (ntaps - 1) / 2.0
    h = np.sinc(2 * 0.4 / M * t) * np.hamming(ntaps)
    taps = (h / h.sum()).astype(np.float32)
    for variant in (0, 1):
        m.tune(decim_variant=variant)
        d = doppler_b200.Decimator(m, taps, M)
        st, sn = None, 0
        for n in (7, 70_001, 150_003):
            buf = (rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype(np.int16) if it == I16 else rng.uniform(-0.7, 0.7, 2 * n).astype(np.float32)).view(np.uint8)
            got, sn = d.mix(buf, it, ot, shift, fs, samplenum=sn)
            want, st = o.mix_decimate(buf, it, ot, shift, fs, taps, M, st)
            ok = sn == st["samplenum"] and np.array_equal(got, want)
            bad += not ok
            print("decimate", "generic" if variant else "register-blocked", M, ntaps, n, "ok" if ok else "MISMATCH")
        d.close()
m.tune(decim_variant=0)
# the resident kernel: a stream of pump blocks with changing shifts, then the context ends while it is resident
sn = sn_ref = 0
for b in range(40):
    buf = rng.integers(-32768, 32768, 2 * 2048, dtype=np.int32).astype(np.int16).view(np.uint8)
    shift = float(rng.uniform(-12000, 12000))
    got, sn = m.mix(buf, I16, I16, shift, 1_024_000, samplenum=sn)
    want, sn_ref = o.mix(buf, I16, I16, shift, 1_024_000, samplenum=sn_ref)
    ok = sn == sn_ref and np.array_equal(got, want)
    bad += not ok
print("resident", "ok" if not bad else "MISMATCH")
m.close()
sys.exit(1 if bad else 0)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > $OUT/san_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|MISMATCH" $OUT/san_$tool.log | head -12
done
