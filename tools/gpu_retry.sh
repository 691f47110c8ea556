#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <file with the command> [extra gpurun flags]: retries while the pod has no free slot (exit code 3)
T=$1; F=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout $T -- "$(cat $F)"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
