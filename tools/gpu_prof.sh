#!/bin/bash
# usage: tools/gpu_prof.sh TAG workload[:param=v,...] ...
cd /root/repo
mkdir -p gpurun_out
tag=$1; shift
for spec in "$@"; do
  w=${spec%%:*}; extra=""
  if [[ "$spec" == *:* ]]; then for kv in $(echo ${spec#*:} | tr ',' ' '); do extra="$extra --param $kv"; done; fi
  name=$(echo $spec | tr ':=,' '___')
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_(div|grad|lift|wave|tp|opmat)' --launch-skip 3 --launch-count 1 -f -o gpurun_out/${tag}_${name} python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/${tag}_${name}.log 2>&1
  tail -2 gpurun_out/${tag}_${name}.log
done
ls -la gpurun_out/${tag}_*.ncu-rep
