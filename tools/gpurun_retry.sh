#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod has no free GPU slot (exit code 3)
log="$1"; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    echo "[retry] attempt $i rc=$rc" >> "$log"
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
