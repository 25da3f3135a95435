#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lift or golden" > gpurun_out/pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest12.log
tail -2 gpurun_out/pytest12.log
for k in lift_fe lift_ef; do timeout 120 python tools/repro.py $k 3999998 2>&1 | tail -1; done
run() { local name=$1 w=$2; shift 2; local extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b12_${name}.json 2> gpurun_out/b12_${name}.err; }
for th in 256 320 352 384; do run lift_$th lift_p4 threads=$th; done
run div_dflt div_p4; run grad_dflt grad_p4; run lift_dflt lift_p4
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b12_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', open(f.replace('.json','.err')).read()[-100:].strip().split('\n')[-1])
PY
bash tools/gpu_prof.sh prof12 div_p4 grad_p4 lift_p4
