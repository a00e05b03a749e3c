#!/bin/bash
# gpurun with retries while the pod answers busy/transient (exit 3); usage: gpurun_retry.sh <log> <timeout> <command...>
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
