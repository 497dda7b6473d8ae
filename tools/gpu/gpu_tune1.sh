#!/bin/bash
# tuning sweep + parity re-check after the kernel template refactor
OUT=gpurun_out/r01c
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== tune"; timeout 900 tools/tune/tune all big > $OUT/tune.jsonl 2> $OUT/tune.err; echo "tune rc=$?"; tail -3 $OUT/tune.err; wc -l $OUT/tune.jsonl
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
