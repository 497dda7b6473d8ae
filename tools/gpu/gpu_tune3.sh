#!/bin/bash
OUT=gpurun_out/r01e
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== tune stream"; timeout 900 tools/tune/tune stream > $OUT/tune_stream.jsonl 2> $OUT/tune.err; echo "tune rc=$?"; tail -3 $OUT/tune.err; wc -l $OUT/tune_stream.jsonl
