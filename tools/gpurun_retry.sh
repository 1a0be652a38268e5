#!/bin/bash
# Local helper (build container): run one gpurun call, retrying while the pod answers "transient" (exit code 3:
# no box / slot free, nothing charged).  usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'
LOG=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun exit $rc (attempt $attempt)" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gave up after 30 transient answers" >> "$LOG"
exit 3
