#!/bin/bash
# A/B of CLI pumps (bin/doppler vs an older build as bin/doppler_old) using the pump's own clock (DOPPLER_STATS)
OUT=gpurun_out/${1:-r01zg}; mkdir -p $OUT
python - <<PY
import numpy as np
rng = np.random.default_rng(1)
rng.integers(-20000, 20000, 2048 * 1024 * 1024, dtype=np.int16).tofile("/dev/shm/iq4g.bin")
PY
export DOPPLER_STATS=1
for rep in 1 2 3; do
for bin in doppler_old doppler; do
  B=doppler_b200/bin/$bin
  [ -x $B ] || continue
  echo -n "{\"bin\": \"$bin\", \"stdin\": \"pipe\", \"stats\": "; cat /dev/shm/iq4g.bin | $B const -s 2000000000 -i i16 --shift -117187500 2>&1 >/dev/null | grep pump_bytes | tr -d '\n'; echo "}"
  echo -n "{\"bin\": \"$bin\", \"stdin\": \"pipe, stdout pipe\", \"stats\": "; (cat /dev/shm/iq4g.bin | $B const -s 2000000000 -i i16 --shift -117187500 2>/tmp/err.txt | cat > /dev/null); grep pump_bytes /tmp/err.txt | tr -d '\n'; echo "}"
done; done | tee $OUT/cli_cmp.jsonl
rm -f /dev/shm/iq4g.bin
