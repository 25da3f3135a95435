#!/bin/bash
# launch list of the default bench command + ncu captures of the kernels without one yet
cd /root/repo
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/b16_plain.json 2> gpurun_out/b16_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu16.log 2>&1
bash tools/gpu_prof.sh prof16 tp_p7 grad_p4_f32 div_p4_f32 lift_p4_f32
for w in tp_p7 grad_p4_f32 div_p4_f32 lift_p4_f32; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/b16_$w.json 2> gpurun_out/b16_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b16_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'GB/s=%.0f'%d['gbs'], 'roof=%.3f'%d['roofline']['roofline_frac'], d['roofline'].get('traffic'))
    except Exception as e:
        print(f, 'ERR', open(f.replace('.json','.err')).read()[-200:].strip().split('\n')[-1])
PY
