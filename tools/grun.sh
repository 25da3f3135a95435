#!/bin/bash
# usage: tools/grun.sh TIMEOUT 'command'   -- gpurun with retries while the pod answers busy (nothing is charged then)
t=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|retry in a few minutes"; then sleep 90; continue; fi
  break
done
