#!/usr/bin/env bash
# gpurun with patience: the pod's GPU slots are shared, a call may come back "transient" (nothing charged). Retry every 60 s.
#   tools/gpurun_retry.sh [--gpus N] --timeout S -- '<command>'
tries=${GPURUN_TRIES:-20}
for i in $(seq 1 "$tries"); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -12
  if echo "$out" | grep -q "status=transient\|backing off\|status=busy"; then
    echo "[retry $i/$tries] no slot; sleeping 60 s" >&2
    sleep 60
    continue
  fi
  exit 0
done
exit 3
