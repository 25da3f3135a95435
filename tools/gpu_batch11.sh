#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "div or grad or lift or golden or unaligned" > gpurun_out/pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest11.log
tail -2 gpurun_out/pytest11.log
for k in grad div lift_fe; do timeout 120 python tools/repro.py $k 3999998 threads=320 2>&1 | tail -1; done
run() { local name=$1 w=$2; shift 2; local extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b11_${name}.json 2> gpurun_out/b11_${name}.err; }
for w in div grad lift; do for th in 256 288 320 352 384 448; do run ${w}_$th ${w}_p4 threads=$th; done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b11_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', open(f.replace('.json','.err')).read()[-100:].strip().split('\n')[-1])
PY
