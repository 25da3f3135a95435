#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
./tools/ubench2 > gpurun_out/ubench2.jsonl 2>&1
cat gpurun_out/ubench2.jsonl
run() { name=$1; w=$2; shift 2; extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b5_${name}.json 2> gpurun_out/b5_${name}.err; }
for th in 128 256; do
  run div_${th}_f2 div_p4 threads=$th flags=2
  run div_${th}_f4 div_p4 threads=$th flags=4
  run div_${th}_f6 div_p4 threads=$th flags=6
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b5_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
