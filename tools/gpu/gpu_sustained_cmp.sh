#!/bin/bash
# interleaved, repeated sustained-load comparison of lean-kernel shapes (the GPU's power / thermal state drifts
# during a long sweep, so single-pass rankings are not reliable)
OUT=gpurun_out/${1:-r01sus}; mkdir -p $OUT
export TUNE_SUSTAINED=1
for rep in 1 2 3 4; do
for v in "i16->i16 W20 S2 U2 smemtab" "i16->i16 W24 S2 U3 smemtab" "i16->i16 W28 S2 U3 smemtab" "i16->i16 W24 S3 U3 smemtab" \
         "i16->f32 W20 S3 U2 smemtab" "i16->f32 W16 S3 U3 smemtab" "i16->f32 W32 S4 U2 smemtab" \
         "f32->i16 W16 S2 U3 smemtab" "f32->i16 W20 S2 U3 smemtab" "f32->i16 W28 S2 U2 smemtab" \
         "f32->f32 W16 S2 U2 smemtab" "f32->f32 W12 S2 U3 smemtab"; do
  tools/tune/tune "stream $v" 2>/dev/null | grep frac
done; done > $OUT/sustained_cmp.jsonl
python - <<PY
import json, collections, statistics
d=collections.defaultdict(list)
for l in open("$OUT/sustained_cmp.jsonl"):
    r=json.loads(l); d[r["variant"]].append(r["gbs"])
for k,v in d.items(): print(k, [round(x) for x in v], "median", round(statistics.median(v)))
PY
