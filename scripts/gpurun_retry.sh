#!/bin/bash
# gpurun with retries on "no box / slot free" (exit 3): scripts/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'
t=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
