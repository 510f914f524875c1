#!/usr/bin/env bash
# usage: tools/gpu_retry.sh <log> <gpurun args...>   — retries while gpurun answers "busy / transient" (exit 3)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "gpurun exit $rc (attempt $i)"; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
