#!/bin/bash
# development aid: retry a gpurun call while the pod answers "transient" (exit code 3 = no box or slot free)
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
