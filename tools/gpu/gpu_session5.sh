#!/bin/bash
# direct-evaluation fast rows: parity, sweep of the table-free cases, ncu of direct periodic
TAG=${1:-r01i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== sweep"; timeout 1200 python tools/sweep.py --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/sweep.log
for c in "direct periodic P=4.9M i16->i16" "direct periodic P=4.9M f32->f32"; do
  n=$(echo "$c" | tr ' =>.' '____' | tr -d '-')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mix_stream -s 3 -c 1 -f -o $OUT/prof_$n \
     python tools/sweep.py --quick --iters 2 --only "$c" --out $OUT/tmp.jsonl > $OUT/ncu_$n.log 2>&1; echo "ncu $c rc=$?"
done
ls -la $OUT
