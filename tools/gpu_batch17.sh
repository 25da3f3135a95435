#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wave3d.py -x -q -k "fp32 or float32" > gpurun_out/pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest17.log
tail -3 gpurun_out/pytest17.log
run() { local name=$1 w=$2; shift 2; local extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b17_${name}.json 2> gpurun_out/b17_${name}.err; }
for w in div grad lift; do for th in 384 512; do run ${w}_f32_$th ${w}_p4_f32 threads=$th; done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b17_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'GB/s=%.0f'%d['gbs'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', open(f.replace('.json','.err')).read()[-200:].strip().split('\n')[-1])
PY
