#!/bin/bash
# tools/gpu/ab_resident.sh -- (GPU box) the per-block host path, variants interleaved on one box: us per block, back to back and
# paced (30-60 us between calls), with blocks of 2048 samples (every request's payload identical: 2048 = 2 periods) and of 2047
# (ragged end, and the plan changes from block to block).  Variant library (when present): tools/tune/libline = an earlier build.
cd "$(dirname "$0")/../.."
run() { PERCALL_QUICK=1 "$@" tools/tune/percall 2> /dev/null | python -c '
import sys, json
r = [json.loads(l) for l in sys.stdin if l.startswith("{")]
print(" ".join("%s%s %.2f" % ("paced-" if "paced" in x else "", "resident" if x["path"].startswith("resident") else "launch", x["us_per_call"]) for x in r))'; }
for round in 1 2 3; do
  for n in 2048 2047; do
    [ -f tools/tune/libline/libdoppler_b200.so ] && echo "n=$n earlier build:   $(run env PERCALL_N=$n LD_LIBRARY_PATH=tools/tune/libline)"
    echo "n=$n product:         $(run env PERCALL_N=$n)"
  done
done
