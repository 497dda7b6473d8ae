#!/bin/bash
# tools/gpu/r02_last.sh [tag] -- evidence run after the resident kernel's protocol change: full GPU suite, bench (+ reference arm),
# sweep, device latency, per-call latency (back to back + paced, phase clock), CLI start-up next to a bare CUDA process, the
# resident / persistent coexistence probe, fuzz, sanitizers.
TAG=${1:-r02last}
bash tools/gpu/session.sh $TAG info smoke test \
  "run:DOPPLER_B200_TRACE=1 tools/tune/percall 2>&1 | tee gpurun_out/$TAG/percall.jsonl | cut -c1-200" \
  "run:bash tools/gpu/cli_startup.sh | tee gpurun_out/$TAG/cli_startup.jsonl" \
  "run:python tools/gpu/resident_coexist.py | tee gpurun_out/$TAG/coexist.txt" \
  "run:python tools/fuzz_parity.py --trials 200 --seed 17 | tee gpurun_out/$TAG/fuzz.txt" \
  bench benchref sweep \
  "run:python tools/latency.py --only device --out gpurun_out/$TAG/latency.jsonl | cut -c1-340" \
  sanitize
python tools/show_bench.py gpurun_out/$TAG/bench.json > gpurun_out/$TAG/bench.txt 2>&1
