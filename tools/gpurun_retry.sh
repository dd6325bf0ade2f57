#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> [--gpus N] -- <command>   (retries while the pod is busy)
log=$1; shift
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then break; fi
  sleep 90
done
echo "finished rc=$rc tries=$try" >> "$log"
