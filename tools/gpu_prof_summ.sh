#!/bin/bash
# usage: tools/gpu_prof_summ.sh TAG workload[@elements][:param=v,...] ...
# like tools/gpu_prof.sh, but the reports are summarised ON the GPU box (tools/ncu_summary.py + tools/ncu_source.py ->
# gpurun_out/TAG_<spec>.txt) and deleted: gpurun merges at most 64 MiB back and one --set full report is ~22 MB
cd /root/repo
mkdir -p gpurun_out
tag=$1; shift
for spec in "$@"; do
  w=${spec%%:*}; extra=""
  if [[ "$spec" == *:* ]]; then for kv in $(echo ${spec#*:} | tr ',' ' '); do extra="$extra --param $kv"; done; fi
  if [[ "$w" == *@* ]]; then extra="$extra --elements ${w#*@}"; w=${w%%@*}; fi
  name=$(echo $spec | tr ':=,@' '____')
  rep=/tmp/${tag}_${name}.ncu-rep
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:'k_(div|grad|lift|wave|tp|opmat|se|hex)' --launch-skip 3 --launch-count 1 -f -o ${rep%.ncu-rep} python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu --no-suite $extra > /tmp/${tag}_${name}.log 2>&1
  {
    echo "# ncu --set full --clock-control none, one launch: bench.py --workload $w $extra (launch 4 of the process)"
    python tools/ncu_summary.py $rep
    python tools/ncu_source.py $rep 12
  } > gpurun_out/${tag}_${name}.txt 2>&1
  rm -f $rep
  head -3 gpurun_out/${tag}_${name}.txt | tail -2 | cut -c1-160
done
