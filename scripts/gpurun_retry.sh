#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: gpurun_retry.sh <timeout_s> <logfile> <command...>
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> $LOG; exit $rc; fi
  sleep 120
done
exit 3
