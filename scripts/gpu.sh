#!/bin/bash
# gpu.sh <log> <timeout> <command...>: run under gpurun, retrying while the pod answers busy (exit 3 / transient)
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1; rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpu.sh done rc=$rc" >> $log
