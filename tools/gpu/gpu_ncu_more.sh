OUT=gpurun_out/r01ncu2; mkdir -p $OUT
for c in "table-L2 P=111145 i16->f32" "table-L2 P=111145 f32->i16" "direct linear (no reset) f32->i16"; do
  n=$(echo "$c" | tr ' =>.()' '______' | tr -d '-')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mix_ -s 3 -c 1 -f -o $OUT/prof_$n \
     python tools/sweep.py --iters 2 --only "$c" --out $OUT/tmp.jsonl > $OUT/ncu_$n.log 2>&1; echo "ncu $c rc=$?"
done
