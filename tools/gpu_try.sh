#!/bin/bash
# retry wrapper: keeps asking for a GPU box until the call is accepted (exit code 3 = no slot, nothing charged)
# usage: tools/gpu_try.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
