#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout_s> <log> <command...>
G=$1; T=$2; LOG=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy\|no box\|retry in a few minutes\|retry later" $LOG && ! grep -q "status=ok" $LOG; then sleep 90; continue; fi
  exit $rc
done
exit 3
