#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab.sh LIB_B workload [workload ...]
cd /root/repo
libb=$1; shift
for w in "$@"; do
  for rep in 1 2 3; do
    for lib in "" "$libb"; do
      FNSM_B200_LIB=$lib python bench.py --workload $w --no-e2e --no-cpu --no-suite --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w', '${lib:-default}', round(l['ms_per_step'], 4), round(l['roofline']['roofline_frac'], 4), l['clocks']['sm_mhz'])"
    done
  done
done
