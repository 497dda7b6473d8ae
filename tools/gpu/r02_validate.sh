#!/bin/bash
# tools/gpu/r02_validate.sh [tag] -- the round-2 validation run: full GPU suite, sweep, latency, per-call, CLI start-up, bench.
TAG=${1:-r02v}
bash tools/gpu/session.sh $TAG info smoke test sweep \
  "run:python tools/latency.py --only device --out gpurun_out/$TAG/latency.jsonl | cut -c1-340" \
  "run:tools/tune/percall | tee gpurun_out/$TAG/percall.jsonl" \
  "run:bash tools/gpu/cli_startup.sh | tee gpurun_out/$TAG/cli_startup.jsonl" \
  bench
python tools/show_bench.py gpurun_out/$TAG/bench.json > gpurun_out/$TAG/bench.txt 2>&1
