#!/bin/bash
# isolated AND sustained timing of a few lean-kernel shapes, interleaved repeats
OUT=gpurun_out/${1:-r01both}; mkdir -p $OUT
for rep in 1 2 3; do
for v in "i16->i16 W20 S2 U2 smemtab" "i16->i16 W16 S2 U3 smemtab" "i16->i16 W20 S2 U3 smemtab" "i16->i16 W16 S3 U3 smemtab" \
         "i16->f32 W20 S3 U2 smemtab" "i16->f32 W16 S2 U3 smemtab" "i16->f32 W20 S2 U3 smemtab" "i16->f32 W12 S3 U3 smemtab"; do
  tools/tune/tune "stream $v" 2>/dev/null | grep frac | sed 's/^{/{"mode": "isolated", /'
  TUNE_SUSTAINED=1 tools/tune/tune "stream $v" 2>/dev/null | grep frac | sed 's/^{/{"mode": "sustained", /'
done; done > $OUT/both_cmp.jsonl
python - <<PY
import json, collections, statistics
d=collections.defaultdict(list)
for l in open("$OUT/both_cmp.jsonl"):
    r=json.loads(l); d[(r["variant"], r["mode"])].append(r["gbs"])
for k,v in sorted(d.items()): print(k[0], k[1], [round(x) for x in v], "median", round(statistics.median(v)))
PY
