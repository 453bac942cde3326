#!/bin/bash
# gpu_retry.sh <timeout-seconds> <command...> -- gpurun with retries while the pod answers "busy" (exit code 3).
# Builder convenience only; not part of the product.
t=$1; shift
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 150
done
exit 3
