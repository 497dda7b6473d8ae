#!/bin/bash
# tools/gpu/session.sh -- ONE parametrised GPU session (replaces the round-1 one-off scripts).
# Usage (repo root, on the GPU box):  bash tools/gpu/session.sh <tag> <step> [<step> ...]
# Steps:
#   info      GPU / host description
#   smoke     __graft_entry__.smoke()
#   test      pytest -m gpu            (TESTS="tests/test_x.py -k expr" narrows it)
#   bench     python bench.py          (BENCH_ARGS="..." appended)
#   benchref  python bench.py --impl reference
#   launches  ncu launch list of the bench command (gpu__time_duration.sum, --clock-control none)
#   ncu:<name>:<kernel regex>:<skip>:<python command...>   ncu --set full of one launch -> prof_<name>.ncu-rep + summary
#   sweep     tools/sweep.py           (SWEEP_ARGS="--only ..." narrows it)
#   sanitize  compute-sanitizer memcheck / racecheck / synccheck (tools/gpu/gpu_sanitize.sh)
#   cli       tools/cli_bench.sh
#   run:<command...>                   any command, logged to run_<n>.log
TAG=${1:?tag}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for step in "$@"; do
  i=$((i+1))
  case "$step" in
    info)
      nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $OUT/gpu.txt 2>&1
      nproc > $OUT/host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/host.txt; ldd --version | head -1 >> $OUT/host.txt; cat $OUT/gpu.txt $OUT/host.txt ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log ;;
    test)
      timeout 2700 python -m pytest ${TESTS:-tests} -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log ;;
    bench)
      t0=$SECONDS; timeout 1200 python bench.py $BENCH_ARGS > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? wall=$((SECONDS - t0)) s"
      cut -c1-400 $OUT/bench.json; tail -3 $OUT/bench.err ;;
    benchref)
      timeout 900 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "benchref rc=$?"; cut -c1-300 $OUT/bench_ref.json ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-ncu --sustained-seconds 0.01 > $OUT/bench_under_ncu.log 2>&1; echo "launch list rc=$?" ;;
    ncu:*)
      IFS=: read -r _ name regex skip cmd <<< "$step"
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o $OUT/prof_$name $cmd > $OUT/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
      ncu -i $OUT/prof_$name.ncu-rep --page raw --csv > $OUT/prof_$name.raw.csv 2>/dev/null && python tools/ncu_summary.py $OUT/prof_$name.raw.csv > $OUT/ncu_$name.txt 2>&1 ;;
    sweep)
      timeout 1500 python tools/sweep.py --out $OUT/sweep.jsonl $SWEEP_ARGS > $OUT/sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/sweep.log ;;
    sanitize)
      bash tools/gpu/gpu_sanitize.sh $TAG ;;
    cli)
      timeout 600 bash tools/cli_bench.sh $OUT > $OUT/cli.log 2>&1; echo "cli rc=$?"; cat $OUT/cli_bench.jsonl ;;
    run:*)
      cmd=${step#run:}
      timeout ${RUN_TIMEOUT:-900} bash -c "$cmd" > $OUT/run_$i.log 2>&1; echo "run[$i] rc=$? : $cmd"; tail -${RUN_TAIL:-25} $OUT/run_$i.log ;;
    *) echo "unknown step $step" ;;
  esac
done
ls $OUT | head -50
