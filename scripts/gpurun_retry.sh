#!/bin/bash
# keeps asking for a GPU slot until the call is actually served (exit code 3 = pod busy, nothing charged)
# usage: scripts/gpurun_retry.sh <timeout_s> [--gpus N] <command...>
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T $G -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
