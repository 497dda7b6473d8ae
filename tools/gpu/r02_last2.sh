#!/bin/bash
# tools/gpu/r02_last2.sh [tag] -- the round's closing run on the final code (after the request moved to tagged sectors and the
# resident kernel fetches its phasors ahead of the input): variants interleaved, full GPU suite, per-call latency with the
# phase clock, bench, fuzz, sanitizers over the resident kernel's tests.  (Sweep / latency / CLI / reference arm: r02_last.sh.)
TAG=${1:-r02last2}
mkdir -p gpurun_out/$TAG
timeout 150 bash tools/gpu/ab_resident.sh | tee gpurun_out/$TAG/ab_resident.txt
bash tools/gpu/session.sh $TAG info smoke test \
  "run:DOPPLER_B200_TRACE=1 tools/tune/percall 2>&1 | tee gpurun_out/$TAG/percall.jsonl | cut -c1-200" \
  "run:python tools/fuzz_parity.py --trials 200 --seed 19 | tee gpurun_out/$TAG/fuzz.txt" \
  bench
for t in memcheck racecheck synccheck initcheck; do
  echo "== $t"; timeout 250 compute-sanitizer --tool $t python -m pytest tests/test_resident.py -x -q -m gpu -k "every_block_size or time_out or plans_beyond or leaves_when" 2>&1 | grep -v "^$" | tail -2
done | tee gpurun_out/$TAG/sanitize_resident.txt
python tools/show_bench.py gpurun_out/$TAG/bench.json > gpurun_out/$TAG/bench.txt 2>&1
