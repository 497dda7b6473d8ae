#!/bin/bash
# round-1 evidence run: tests, bench, launch list, ncu --set full of the kernels DESIGN.md names, sweep, CLI
TAG=${1:-r01u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/host.txt; ldd --version | head -1 >> $OUT/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -3 $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --seconds 16 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (bench kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mix_grid -s 3 -c 1 -f -o $OUT/prof_bench_f32_i16 \
    python bench.py --steps 2 --warmup 3 --seconds 16 --no-cpu-baseline --no-e2e > $OUT/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
for c in "table-L2 P=111145 i16->i16" "table-L2 P=111145 f32->f32" "table-smem P=256 i16->i16" "direct linear (no reset) i16->i16"; do
  n=$(echo "$c" | tr ' =>.()' '______' | tr -d '-')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mix_ -s 3 -c 1 -f -o $OUT/prof_$n \
     python tools/sweep.py --iters 2 --only "$c" --out $OUT/tmp.jsonl > $OUT/ncu_$n.log 2>&1; echo "ncu $c rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mix_stream -s 2 -c 1 -f -o $OUT/prof_cfg3 \
    python tools/sweep.py --iters 4 --only "cfg3 track" --out $OUT/tmp.jsonl > $OUT/ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
echo "== sweep"; timeout 1200 python tools/sweep.py --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/sweep.log
echo "== cli"; timeout 600 bash tools/cli_bench.sh $OUT > $OUT/cli.log 2>&1; echo "cli rc=$?"; cat $OUT/cli_bench.jsonl
ls -la $OUT
