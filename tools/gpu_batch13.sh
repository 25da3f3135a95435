#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wave3d.py -x -q > gpurun_out/pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest13.log
tail -3 gpurun_out/pytest13.log
( time python bench.py ) > gpurun_out/b13_default.json 2> gpurun_out/b13_default.err; tail -3 gpurun_out/b13_default.err
( time python bench.py --impl reference ) > gpurun_out/b13_reference.json 2> gpurun_out/b13_reference.err; tail -3 gpurun_out/b13_reference.err
timeout 600 python bench.py --workload wave_p4 --steps 10 --warmup 3 > gpurun_out/b13_wave.json 2> gpurun_out/b13_wave.err
timeout 600 python bench.py --workload wave_p4_f32 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/b13_wave_f32.json 2> gpurun_out/b13_wave_f32.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b13_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'roof=%s'%r.get('roofline_frac'), 'e2e=%s'%((d.get('e2e') or {}).get('value')), 'cpu=%s'%((d.get('cpu_baseline') or {}).get('value')))
    except Exception as e:
        print(f, 'ERR', open(f.replace('.json','.err')).read()[-300:])
PY
