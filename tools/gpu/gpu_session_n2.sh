#!/bin/bash
# N=2 bench (both arms), launched the way the driver launches it
OUT=gpurun_out/r01n2
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "n2 rc=$?"; cat $OUT/bench_n2.json; tail -3 $OUT/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/ref_n2.json 2> $OUT/ref_n2.err; echo "ref n2 rc=$?"; cat $OUT/ref_n2.json; tail -3 $OUT/ref_n2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/ref_n1.json 2> $OUT/ref_n1.err; echo "ref n1 rc=$?"; cat $OUT/ref_n1.json; tail -3 $OUT/ref_n1.err
