#!/bin/bash
# tools/gpu/r02_evidence.sh [tag] -- the round-2 evidence run on one GPU: tests, bench (+ reference arm), launch list, ncu captures of
# the kernels DESIGN.md names, sweep, latency, per-call, CLI, A/B, sanitizers, fuzz.
TAG=${1:-r02z}
bash tools/gpu/session.sh $TAG info smoke test bench benchref launches \
  "ncu:bench_f32_i16:mix_grid:3:python tools/ncu_traffic.py --samples 640000000 --launches 5" \
  "ncu:column_i16_i16:mix_stream:3:python tools/sweep.py --iters 2 --only table-L2_P=111145_i16->i16 --out gpurun_out/$TAG/tmp.jsonl" \
  "ncu:cfg3:mix_stream:2:python tools/sweep.py --iters 4 --only cfg3_track --out gpurun_out/$TAG/tmp.jsonl" \
  "ncu:direct_linear_i16_i16:mix_grid:3:python tools/sweep.py --iters 2 --only direct_linear_(no_reset)_i16->i16 --out gpurun_out/$TAG/tmp.jsonl" \
  "ncu:small_256k:mix_small:3:python tools/sweep.py --iters 2 --only cfg5_1s_@_256000 --out gpurun_out/$TAG/tmp.jsonl" \
  sweep \
  "run:python tools/latency.py --out gpurun_out/$TAG/latency.jsonl | cut -c1-200" \
  "run:DOPPLER_B200_TRACE=1 tools/tune/percall 2>&1 | tee gpurun_out/$TAG/percall.jsonl" \
  "run:bash tools/gpu/cli_startup.sh | tee gpurun_out/$TAG/cli_startup.jsonl" \
  "run:python tools/ab_seg.py gpurun_out/$TAG/ab_seg.jsonl" \
  "run:python tools/fuzz_parity.py --trials 160 --seed 7 | tee gpurun_out/$TAG/fuzz.txt" \
  sanitize cli
python tools/show_bench.py gpurun_out/$TAG/bench.json > gpurun_out/$TAG/bench.txt 2>&1
