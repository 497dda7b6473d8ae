#!/bin/bash
# CLI throughput: `doppler const` stdin -> stdout through pipes, input from tmpfs
OUT=${1:-gpurun_out/cli}
mkdir -p $OUT
BIN=doppler_b200/bin/doppler
F=/dev/shm/iq_i16.bin
python - <<PY
import numpy as np
rng = np.random.default_rng(1)
a = rng.integers(-20000, 20000, 512 * 1024 * 1024, dtype=np.int16)   # 1 GiB = 256 Mi samples i16
a.tofile("$F")
PY
for pair in "i16 i16" "i16 f32"; do
  set -- $pair
  for rep in 1 2 3; do
    s=$(date +%s.%N)
    cat $F | $BIN const -s 2000000000 -i $1 -o $2 --shift -117187500 2>/dev/null | cat > /dev/null
    e=$(date +%s.%N)
    python -c "n=268435456; t=$e-$s; print('{\"cli\": \"const $1->$2 via pipes\", \"samples\": %d, \"seconds\": %.3f, \"msps\": %.1f, \"in_MBps\": %.0f}' % (n, t, n/t/1e6, n*4/t/1e6))"
  done
done | tee $OUT/cli_bench.jsonl
# the pipe alone, for scale
s=$(date +%s.%N); cat $F | cat > /dev/null; e=$(date +%s.%N)
python -c "t=$e-$s; print('{\"cli\": \"cat | cat (pipe ceiling)\", \"seconds\": %.3f, \"in_MBps\": %.0f}' % (t, 1073.74/t))" | tee -a $OUT/cli_bench.jsonl
rm -f $F
