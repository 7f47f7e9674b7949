#!/bin/bash
# local helper: gpurun with retries while the pod answers "transient/busy" (nothing charged in that case)
# usage: gpurun_retry.sh <timeout_s> '<command>' [extra gpurun flags]
T=$1; shift; CMD=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T "$@" -- "$CMD" 2>&1)
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|no box or slot"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
