#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
run() { local name=$1 w=$2; shift 2; local extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b7_${name}.json 2> gpurun_out/b7_${name}.err; }
for f in 0 2 4 6 8; do run grad_256_f$f grad_p4 threads=256 flags=$f; done
for f in 8; do run div_256_f$f div_p4 threads=256 flags=$f; done
run grad_128_f0 grad_p4 threads=128
run grad_128_f6 grad_p4 threads=128 flags=6
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b7_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-200:])
PY
