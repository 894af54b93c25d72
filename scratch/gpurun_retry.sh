#!/bin/bash
# gpurun with retries while the pod answers "no slot free" (exit code 3; nothing is charged for those)
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3
