#!/bin/bash
# development aid: retry a gpurun call while the pod answers "busy" (exit code 3: nothing charged)
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc attempt=$i" >> $log; exit $rc; fi
  sleep 150
done
echo "gave up" >> $log
