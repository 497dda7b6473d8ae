#!/bin/bash
# full GPU suite (incl. full-size tests), bench, sweep
TAG=${1:-r01m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -16 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json | cut -c1-400; tail -3 $OUT/bench.err
echo "== sweep"; timeout 1200 python tools/sweep.py --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/sweep.log
