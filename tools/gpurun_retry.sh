#!/bin/bash
# usage: tools/gpurun_retry.sh LOG [gpurun args...] -- retries while the pod answers busy/transient (nothing charged)
log=$1; shift
for i in $(seq 1 12); do
  gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|rc=None" "$log"; then sleep 200; else break; fi
done
