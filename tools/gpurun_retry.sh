#!/bin/bash
# tools/gpurun_retry.sh [gpurun options] -- '<command>': retry while the pod answers "busy" (exit code 3, nothing charged)
for try in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] busy, try $try; sleeping 90 s"
  sleep 90
done
exit 3
