#!/bin/bash
# Retry a gpurun call while the pod answers "busy / transient" (exit code 3); all arguments go to gpurun.
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $i: pod busy, sleeping 150 s"
  sleep 150
done
exit 3
