#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wave3d.py tests/test_gpu_measure.py -x -q -k "fp32 or float32" > gpurun_out/pytest18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest18.log
tail -12 gpurun_out/pytest18.log
for w in div grad lift; do for th in 384 512; do
  python bench.py --workload ${w}_p4_f32 --steps 10 --warmup 3 --no-e2e --no-cpu --param threads=$th | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', $th, d['ms_per_step'], d['roofline']['roofline_frac'])"
done; done
python tools/report.py > gpurun_out/report.md 2> gpurun_out/report.err; cat gpurun_out/report.md; tail -3 gpurun_out/report.err
