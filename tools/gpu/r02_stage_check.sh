#!/bin/bash
# tools/gpu/r02_stage_check.sh -- (GPU box) the resident kernel after a change: its tests, the per-block latency with a shift that
# has a phasor table (5000 Hz) and one that has none (7321.7 Hz) against an earlier build (tools/tune/libline, when present),
# fuzz over the per-block path, memcheck / racecheck over the resident kernel's tests.
q() { PERCALL_QUICK=1 "$@" tools/tune/percall 2>/dev/null | python -c '
import sys, json
r = [json.loads(l) for l in sys.stdin if l.startswith("{")]
print(" ".join("%s%s %.2f" % ("paced-" if "paced" in x else "", "resident" if x["path"].startswith("resident") else "launch", x["us_per_call"]) for x in r))'; }
timeout 120 python -m pytest tests/test_resident.py -x -q -m gpu 2>&1 | tail -2
for sh in 5000 7321.7; do
  [ -f tools/tune/libline/libdoppler_b200.so ] && echo "shift $sh previous build: $(q env PERCALL_SHIFT=$sh LD_LIBRARY_PATH=tools/tune/libline)"
  echo "shift $sh product:        $(q env PERCALL_SHIFT=$sh)"
done
timeout 60 python tools/fuzz_parity.py --trials 60 --seed 59 | tail -1
for t in memcheck racecheck; do timeout 120 compute-sanitizer --tool $t python -m pytest tests/test_resident.py -x -q -m gpu -k "every_block_size or time_out or plans_beyond" 2>&1 | grep -v "^$" | tail -2; done
