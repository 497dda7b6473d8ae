#!/bin/bash
# tools/gpu/ab_resident.sh -- (GPU box) the per-block host path, variants interleaved on one box: us per 8192-byte block,
# back to back and paced (30-60 us between calls).  Variant library (when present): tools/tune/libhead.
cd "$(dirname "$0")/../.."
run() { PERCALL_QUICK=1 DOPPLER_B200_TRACE=1 "$@" tools/tune/percall 2> /tmp/percall_trace.txt | python -c '
import sys, json
r = [json.loads(l) for l in sys.stdin if l.startswith("{")]
print(" ".join("%s%s %.2f%s" % ("paced-" if "paced" in x else "", "resident" if x["path"].startswith("resident") else "launch", x["us_per_call"], " (median %.2f)" % x["median_us"] if "paced" in x else "") for x in r))'; grep resident /tmp/percall_trace.txt | cut -c1-200; }
for round in 1 2 3; do
  [ -f tools/tune/libhead/libdoppler_b200.so ] && echo "head: $(run env LD_LIBRARY_PATH=tools/tune/libhead)"
  echo "new: $(run env)"
  echo "new, planner on every block: $(run env DOPPLER_B200_NO_STEADY_RULE=1)"
done
