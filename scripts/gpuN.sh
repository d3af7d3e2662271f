#!/bin/bash
# gpuN.sh <N> <log> <timeout> <command...>: like gpu.sh with --gpus N
n=$1; log=$2; to=$3; shift 3
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1; rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpuN.sh done rc=$rc" >> $log
