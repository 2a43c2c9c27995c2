#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
# usage: [GPUS=N] gpurun_retry.sh <timeout> '<command>'
for i in $(seq 1 40); do
    if [ -n "$GPUS" ]; then
        /usr/local/graft/bin/gpurun --gpus "$GPUS" --timeout "$1" -- "$2"
    else
        /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
    fi
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 45
done
exit 3
