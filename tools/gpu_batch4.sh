#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
run() { # name workload params...
  name=$1; w=$2; shift 2; extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b4_${name}.json 2> gpurun_out/b4_${name}.err
}
for st in 0 500 1000 1500 2500; do
  run div_256_s$st div_p4 threads=256 stagger=$st
  run grad_256_s$st grad_p4 threads=256 stagger=$st
  run lift_256_s$st lift_p4 threads=256 stagger=$st
done
for st in 400 800 1600; do
  run lift_384_s$st lift_p4 threads=384 stagger=$st
  run grad_320_s$st grad_p4 threads=320 stagger=$st
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'GB/s=%.0f'%d['gbs'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
