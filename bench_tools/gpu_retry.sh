#!/bin/bash
# gpu_retry.sh [--gpus N] TIMEOUT 'command' : gpurun with retries while the pod answers "busy" (exit 3)
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
