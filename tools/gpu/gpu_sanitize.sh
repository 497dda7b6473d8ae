#!/bin/bash
# compute-sanitizer over the kernels' paths (small inputs): memcheck, racecheck (shared-memory hazards), synccheck
OUT=gpurun_out/${1:-r01v}
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
    for shift, fs, n in [(-15000.0, 256000, 70_001), (-3_912_345.25, 200_000_000, 200_003), (7321.7, 1_024_000, 130_001),
                         (1.0, 2_000_000_000, 50_001)]:
        if it == I16:
            buf = rng.integers(-32768, 32768, 2 * n, dtype=np.int32).astype(np.int16).view(np.uint8)
        else:
            buf = rng.uniform(-0.7, 0.7, 2 * n).astype(np.float32).view(np.uint8)
        got, sn = m.mix(buf, it, ot, shift, fs)
        want, sn_ref = o.mix(buf, it, ot, shift, fs)
        ok = sn == sn_ref and np.array_equal(got, want)
        bad += not ok
        print(it, ot, shift, n, "ok" if ok else "MISMATCH")
m.close()
sys.exit(1 if bad else 0)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > $OUT/san_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|MISMATCH" $OUT/san_$tool.log | head -12
done
echo "== cli timing"
BIN=doppler_b200/bin/doppler
python - <<PY
import numpy as np
rng = np.random.default_rng(1)
rng.integers(-20000, 20000, 2048 * 1024 * 1024, dtype=np.int16).tofile("/dev/shm/iq4g.bin")   # 4 GiB = 1 Gi samples
PY
s=$(date +%s.%N); $BIN const -s 2000000000 -i i16 --shift -117187500 < /dev/null > /dev/null 2>&1; e=$(date +%s.%N)
python -c "print('{\"cli\": \"empty input (start-up)\", \"seconds\": %.3f}' % ($e-$s))" | tee $OUT/cli2.jsonl
for mode in pipe file; do
  for rep in 1 2; do
    s=$(date +%s.%N)
    if [ $mode = pipe ]; then cat /dev/shm/iq4g.bin | $BIN const -s 2000000000 -i i16 --shift -117187500 2>/dev/null | cat > /dev/null
    else $BIN const -s 2000000000 -i i16 --shift -117187500 < /dev/shm/iq4g.bin > /dev/null 2>/dev/null; fi
    e=$(date +%s.%N)
    python -c "n=1073741824; t=$e-$s; print('{\"cli\": \"const i16->i16 4 GiB, stdin=$mode\", \"seconds\": %.3f, \"msps\": %.1f, \"in_MBps\": %.0f}' % (t, n/t/1e6, n*4/t/1e6))" | tee -a $OUT/cli2.jsonl
  done
done
rm -f /dev/shm/iq4g.bin
