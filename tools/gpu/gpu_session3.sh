#!/bin/bash
# one gpurun call: smoke, GPU parity tests, bench, sweep, ncu launch list + full captures
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/host.txt; ldd --version | head -1 >> $OUT/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== sweep"; timeout 1200 python tools/sweep.py --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/sweep.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --seconds 16 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (mix kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mix_stream -s 3 -c 1 -f -o $OUT/prof_mix_f32_i16 \
    python bench.py --steps 2 --warmup 3 --seconds 16 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
for c in "table-smem P=256 i16->i16" "table-L2 P=111145 i16->i16" "direct periodic P=4.9M i16->i16"; do
  n=$(echo "$c" | tr ' =>.' '____' | tr -d '-')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mix_stream -s 3 -c 1 -f -o $OUT/prof_$n \
     python tools/sweep.py --quick --iters 2 --only "$c" --out $OUT/tmp.jsonl > $OUT/ncu_$n.log 2>&1; echo "ncu $c rc=$?"
done
ls -la $OUT
