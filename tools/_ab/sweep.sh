#!/bin/bash
cd /root/repo
for rep in 1 2; do
for v in "$@"; do
  for w in ${WL:-div_p4 se_p4}; do
  FNSM_B200_LIB=/root/repo/tools/_ab/libv_$v.so timeout 100 python bench.py --workload $w --no-e2e --no-cpu --no-suite --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w', '$v', round(l['ms_per_step'], 4), round(l['roofline']['roofline_frac'], 4))"
  done
done
done
