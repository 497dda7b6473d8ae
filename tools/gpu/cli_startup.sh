#!/bin/bash
# tools/gpu/cli_startup.sh -- wall clock of the CLI on BASELINE configs[0] (1 s of i16 IQ @ 256 ksps = 1 MB on stdin):
# process start to exit, with the pump's own start-up clock (DOPPLER_STATS).
python -c "import numpy as np; np.random.default_rng(1).integers(-20000,20000,512000,dtype=np.int16).tofile('/tmp/x.iq')"
for i in 1 2 3 4; do
  s=$(date +%s%N)
  DOPPLER_STATS=1 doppler_b200/bin/doppler const -s 256000 -i i16 --shift -15000 < /tmp/x.iq 2>/tmp/err.txt >/tmp/y.iq
  e=$(date +%s%N)
  echo "{\"run\": $i, \"wall_ms\": $(( (e-s)/1000000 )), \"stats\": [$(grep -E '^\{' /tmp/err.txt | paste -sd, -)]}"
done
# the floor: a process that does nothing but create a CUDA context (tools/tune/ctxtime.cu)
for i in 1 2 3; do
  s=$(date +%s%N); o=$(tools/tune/ctxtime); e=$(date +%s%N)
  echo "{\"run\": $i, \"wall_ms\": $(( (e-s)/1000000 )), \"stats\": [$o]}"
done
