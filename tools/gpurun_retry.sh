#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers "busy / draining" (nothing is charged then)
out=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if grep -q "status=transient" "$out"; then sleep 150; continue; fi
  break
done
tail -40 "$out"
