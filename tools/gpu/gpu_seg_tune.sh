#!/bin/bash
# segmented-kernel shape sweep: product library (SEGV 0) and the `make variants` builds
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
ONLY="P=111145,cfg3 track,cfg4 track"
echo "== product sweep"; timeout 900 python tools/sweep.py --only "$ONLY" --out $OUT/sweep_v0.jsonl > $OUT/sweep_v0.log 2>&1; echo "rc=$?"; cat $OUT/sweep_v0.log
for v in 1 2 3; do
  export DOPPLER_B200_LIB=doppler_b200/libdoppler_b200_segv$v.so
  echo "== variant $v parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "column or direct or block_schedule or replay" > $OUT/pytest_v$v.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest_v$v.log
  echo "== variant $v sweep"; timeout 900 python tools/sweep.py --only "$ONLY" --out $OUT/sweep_v$v.jsonl > $OUT/sweep_v$v.log 2>&1; echo "rc=$?"; cat $OUT/sweep_v$v.log
  unset DOPPLER_B200_LIB
done
