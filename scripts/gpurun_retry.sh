#!/bin/bash
# usage: scripts/gpurun_retry.sh <tag> <timeout_s> [--gpus N] -- '<command>'
# retries while gpurun answers 3 (no box / slot free, nothing charged); log in gpurun_out/call_<tag>.log
tag=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
mkdir -p gpurun_out
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$1" > "gpurun_out/call_${tag}.log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
