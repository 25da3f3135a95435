#!/bin/bash
# first GPU batch of the session: parity + large-size cross-checks + benches per config
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for k in grad div lift_fe lift_ef; do
  for n in 300000 2000000 4000000; do
    for th in 256 384; do
      timeout 120 python tools/repro.py $k $n threads=$th > gpurun_out/repro_${k}_${n}_${th}.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/repro_${k}_${n}_${th}.log
    done
  done
done
for w in div_p4 grad_p4 lift_p4; do
  for th in 128 256 384; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu --param threads=$th > gpurun_out/b1_${w}_${th}.json 2> gpurun_out/b1_${w}_${th}.err
  done
done
timeout 300 python bench.py --workload tp_p7 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/b1_tp_p7.json 2> gpurun_out/b1_tp_p7.err
tail -3 gpurun_out/pytest_gpu.log
grep -h "max rel\|FAILED\|Error" gpurun_out/repro_*.log | sort | uniq -c | head -40
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b1_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'GF=%.0f'%d['value'], 'GB/s=%.0f'%d['gbs'], 'roof=%.3f'%d['roofline']['roofline_frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
