#!/bin/bash
# Retries a gpurun call while the pod answers "transient/busy" (nothing is charged for those).
#   tools/gpurun_retry.sh <timeout_s> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for attempt in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$CMD" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|status=busy\|exit code 3\|rc=3"; then
    echo "[retry $attempt] pod busy, sleeping 90 s" >&2; sleep 90; continue
  fi
  echo "$OUT"; exit 0
done
echo "gave up after 40 attempts"; exit 3
