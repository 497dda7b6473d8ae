OUT=gpurun_out/r01iso; mkdir -p $OUT
for rep in 1 2 3; do
for v in "i16->i16 W20 S2 U2 smemtab" "i16->i16 W24 S2 U3 smemtab" "i16->i16 W28 S2 U3 smemtab" "i16->f32 W20 S3 U2 smemtab" "i16->f32 W16 S3 U3 smemtab" "i16->f32 W24 S4 U2 smemtab"; do
  tools/tune/tune "stream $v" 2>/dev/null | grep frac
done; done > $OUT/isolated_cmp.jsonl
python - <<PY
import json, collections, statistics
d=collections.defaultdict(list)
for l in open("$OUT/isolated_cmp.jsonl"):
    r=json.loads(l); d[r["variant"]].append(r["gbs"])
for k,v in d.items(): print(k, [round(x) for x in v], "median", round(statistics.median(v)))
PY
