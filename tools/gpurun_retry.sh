#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   - retries while the pod answers "busy" (exit 3)
t=$1; shift
for i in $(seq 1 20); do
    /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 90
done
exit 3
