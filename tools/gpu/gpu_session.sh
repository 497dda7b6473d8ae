#!/bin/bash
# tools/gpu/gpu_session.sh -- one gpurun call: smoke, GPU parity tests, bench, ncu launch list + full capture.
# Usage (from the repo root, on the GPU box): bash tools/gpu/gpu_session.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/host.txt; ldd --version | head -1 >> $OUT/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --seconds 16 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (mix kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mix_kernel -s 3 -c 2 -f -o $OUT/prof_mix \
    python bench.py --steps 2 --warmup 3 --seconds 16 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
