#!/bin/bash
# usage: tools/_gpurun_retry.sh <timeout_s> <logname> -- retries gpurun while the pod answers busy (exit 3)
T=$1; L=$2
for i in $(seq 1 20); do
  timeout $((T + 2000)) gpurun --timeout $T -- "bash tools/_gpu1.sh > gpurun_out/$L 2>&1; tail -60 gpurun_out/$L"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
