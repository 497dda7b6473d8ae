#!/bin/bash
# re-entry check: smoke, GPU parity tests, bench, instruction-cost microbenchmark, big-buffer probe
TAG=${1:-r01h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/host.txt; ldd --version | head -1 >> $OUT/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pipes"; timeout 300 tools/tune/pipes > $OUT/pipes.jsonl 2> $OUT/pipes.err; echo "pipes rc=$?"; cat $OUT/pipes.jsonl
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bigbuf"; timeout 600 python tools/bigbuf_probe.py > $OUT/bigbuf.log 2>&1; echo "bigbuf rc=$?"; tail -50 $OUT/bigbuf.log
