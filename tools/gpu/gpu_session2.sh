#!/bin/bash
# bench (fixed stream), full sweep, ncu captures of the direct and L2-table variants
TAG=${1:-r01b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== sweep"; timeout 1200 python tools/sweep.py --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/sweep.log
for c in "direct periodic P=4.9M i16->i16" "table-L2 P=111145 i16->i16" "table-smem P=256 i16->i16" "direct periodic P=4.9M f32->f32"; do
  n=$(echo "$c" | tr ' =>.' '____' | tr -d '-')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mix_kernel -s 3 -c 1 -f -o $OUT/prof_$n \
     python tools/sweep.py --quick --iters 2 --only "$c" --out $OUT/tmp.jsonl > $OUT/ncu_$n.log 2>&1; echo "ncu $c rc=$?"
done
ls -la $OUT
