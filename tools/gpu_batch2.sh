#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for w in div_p4 grad_p4 lift_p4; do
  for th in 128 256 384; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu --param threads=$th > gpurun_out/b2_${w}_${th}.json 2> gpurun_out/b2_${w}_${th}.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'GB/s=%.0f'%d['gbs'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-500:])
PY
