#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "grad or golden" > gpurun_out/pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest8.log
tail -2 gpurun_out/pytest8.log
timeout 120 python tools/repro.py grad 4000000 threads=256 2>&1 | tail -1
run() { local name=$1 w=$2; shift 2; local extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b8_${name}.json 2> gpurun_out/b8_${name}.err; }
for f in 0 2 4 6; do run grad_256_f$f grad_p4 threads=256 flags=$f; done
run grad_320_f0 grad_p4 threads=320
run grad_128_f0 grad_p4 threads=128
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-200:])
PY
