#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged):  tools/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc attempt=$i" >> "$LOG"; exit $rc; fi
  sleep 120
done
echo "gave up" >> "$LOG"; exit 3
