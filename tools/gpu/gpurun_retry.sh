#!/bin/bash
# tools/gpu/gpurun_retry.sh LOG [gpurun args...] -- (authoring container) retries a gpurun call while the pod answers "transient".
LOG=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG" || grep -q "exit code 3" "$LOG"; then sleep 90; continue; fi
  break
done
