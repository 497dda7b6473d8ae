#!/bin/bash
# tools/gpu/r02_multi.sh N [tag] -- the N-GPU runs: bench.py under torchrun as the driver launches it, the device-group tests, and the
# one-process multi-GPU bench (doppler_b200_multi_*).
N=${1:-2}; TAG=${2:-r02m$N}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/gpus.txt; nproc >> $OUT/gpus.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
   > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N rc=$?"; cut -c1-260 $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
python tools/show_bench.py $OUT/bench_n$N.json > $OUT/bench_n$N.txt 2>&1
timeout 600 python bench.py --impl reference --gpus $N > $OUT/bench_ref_n$N.json 2>/dev/null; echo "ref rc=$?"
timeout 900 python -m pytest tests/test_multi.py -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -2 $OUT/pytest_multi.log
timeout 900 python tools/multi_bench.py > $OUT/multi_bench.json 2> $OUT/multi_bench.err; echo "multi_bench rc=$?"; cat $OUT/multi_bench.json; tail -3 $OUT/multi_bench.err
