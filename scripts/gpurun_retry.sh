#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <command string>   -- retries while the pod answers "transient" (no box free)
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|no box"; then
    sleep 90
    continue
  fi
  echo "$out"
  exit 0
done
echo "gave up after 40 tries"; echo "$out" | tail -5
