#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "div or grad or lift or golden or unaligned" > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest3.log
tail -3 gpurun_out/pytest3.log
for k in grad div lift_fe; do
  timeout 120 python tools/repro.py $k 4000000 threads=256 > gpurun_out/repro3_${k}.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/repro3_${k}.log
  tail -2 gpurun_out/repro3_${k}.log
done
for w in div_p4 grad_p4 lift_p4; do
  for th in 128 256 320 384; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu --param threads=$th > gpurun_out/b3_${w}_${th}.json 2> gpurun_out/b3_${w}_${th}.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'GB/s=%.0f'%d['gbs'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
for w in div_p4 grad_p4 lift_p4; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:dmma --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof3_${w} python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu3_${w}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
