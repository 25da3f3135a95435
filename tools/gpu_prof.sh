#!/bin/bash
# usage: tools/gpu_prof.sh TAG workload[@elements][:param=v,...] ...   -> gpurun_out/TAG_<spec>.ncu-rep (+ .log)
cd /root/repo
mkdir -p gpurun_out
tag=$1; shift
for spec in "$@"; do
  w=${spec%%:*}; extra=""
  if [[ "$spec" == *:* ]]; then for kv in $(echo ${spec#*:} | tr ',' ' '); do extra="$extra --param $kv"; done; fi
  if [[ "$w" == *@* ]]; then extra="$extra --elements ${w#*@}"; w=${w%%@*}; fi
  name=$(echo $spec | tr ':=,@' '____')
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_(div|grad|lift|wave|tp|opmat|se|hex)' --launch-skip 3 --launch-count 1 -f -o gpurun_out/${tag}_${name} python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu --no-suite $extra > gpurun_out/${tag}_${name}.log 2>&1
  tail -2 gpurun_out/${tag}_${name}.log | cut -c1-300
done
ls -la gpurun_out/${tag}_*.ncu-rep
