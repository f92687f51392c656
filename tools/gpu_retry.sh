#!/bin/bash
# usage: tools/gpu_retry.sh <out-prefix> <gpus> <timeout> <command...>   — retries while the pod answers "busy" (exit code 3)
P=$1; G=$2; T=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > gpurun_out/$P.stdout 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "done rc=$rc after $i tries"; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
