#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3, nothing charged).  Usage: tools/gpurun_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 20); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] busy, attempt $i; sleeping 90 s"
    sleep 90
done
exit 3
