#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <log> <command...>   (retries while the pod answers busy / transient)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy\|no box\|retry in a few minutes" $LOG && ! grep -q "status=ok" $LOG; then sleep 60; continue; fi
  exit $rc
done
exit 3
