#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <tag> '<command>'   -- retries while the pod answers "busy" (exit 3)
T=$1; TAG=$2; CMD=$3
for i in $(seq 1 12); do
    /usr/local/graft/bin/gpurun --timeout $T -- "$CMD" > gpurun_out/${TAG}_call.log 2>&1
    rc=$?
    echo "exit $rc (try $i)" >> gpurun_out/${TAG}_call.log
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 100
done
exit 3
