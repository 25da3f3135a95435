#!/bin/bash
# alignment-cliff numbers: E = 4 000 001 (odd) / 4 000 002 (even, not a multiple of 4) for every p = 4 DG kernel
cd /root/repo; mkdir -p gpurun_out
out=gpurun_out/${1:-cliff}.jsonl; : > $out
for w in grad_p4 div_p4 lift_p4 grad_p4_f32 div_p4_f32 lift_p4_f32; do
  for e in 4000000 4000001 4000002; do
    python bench.py --workload $w --elements $e --no-e2e --no-cpu --no-suite --steps 10 --warmup 3 2>>gpurun_out/cliff.err | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'workload': '$w', 'E': $e, 'ms': round(l['ms_per_step'], 4), 'gflops': round(l['value']), 'gbs': round(l['gbs']), 'frac': round(l['roofline']['roofline_frac'], 3)}))" | tee -a $out
  done
done
