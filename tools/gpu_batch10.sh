#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest10.log
tail -4 gpurun_out/pytest10.log
for k in grad div lift_fe lift_ef; do for n in 4000000 3999998 1000010; do
  timeout 120 python tools/repro.py $k $n threads=256 2>&1 | tail -1
done; done
run() { local name=$1 w=$2; shift 2; local extra=""; for kv in "$@"; do extra="$extra --param $kv"; done
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu $extra > gpurun_out/b10_${name}.json 2> gpurun_out/b10_${name}.err; }
for w in div grad lift; do for th in 128 256 320 384; do run ${w}_$th ${w}_p4 threads=$th; done; done
for w in div grad; do run ${w}_256_f6 ${w}_p4 threads=256 flags=6; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b10_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', open(f.replace('.json','.err')).read()[-150:])
PY
