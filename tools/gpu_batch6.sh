#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "div or grad or lift or golden or unaligned" > gpurun_out/pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest6.log
tail -2 gpurun_out/pytest6.log
run() { local name=$1 w=$2; shift 2; local extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b6_${name}.json 2> gpurun_out/b6_${name}.err; }
for w in div grad lift; do
  for th in 256 320 384; do run ${w}_$th ${w}_p4 threads=$th; done
done
run div_256_f6 div_p4 threads=256 flags=6
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b6_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-200:])
PY
