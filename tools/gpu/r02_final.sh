#!/bin/bash
# tools/gpu/r02_final.sh [tag] -- the end-of-round evidence run on one GPU after the register-blocked decimator and the resident
# kernel: tests, bench (+ reference arm), launch list, ncu of the bench kernel and of the decimator, per-call latency, fuzz,
# sanitizers.  (Sweeps / latency / CLI / A-B: tools/gpu/r02_evidence.sh.)
TAG=${1:-r02f}
bash tools/gpu/session.sh $TAG info smoke test bench benchref launches \
  "ncu:bench_f32_i16:mix_grid:3:python tools/ncu_traffic.py --samples 640000000 --launches 5" \
  "run:python tools/decim_bench.py gpurun_out/$TAG/decim_bench.jsonl | cut -c1-400" \
  "ncu:decim_f32_i16:mix_decimate_fast:1:env DECIM_CASES=0 python tools/decim_bench.py gpurun_out/$TAG/tmp.jsonl 128000000" \
  "ncu:direct_linear_i16_i16:mix_grid:3:python tools/sweep.py --iters 2 --only direct_linear_(no_reset)_i16->i16 --out gpurun_out/$TAG/tmp.jsonl" \
  sweep \
  "run:DOPPLER_B200_TRACE=1 tools/tune/percall 2>&1 | tee gpurun_out/$TAG/percall.jsonl" \
  "run:python tools/fuzz_parity.py --trials 200 --seed 11 | tee gpurun_out/$TAG/fuzz.txt" \
  sanitize
python tools/show_bench.py gpurun_out/$TAG/bench.json > gpurun_out/$TAG/bench.txt 2>&1
