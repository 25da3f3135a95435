python -m pytest tests -m gpu -x -q -k "div or wave" 2>&1 | tail -3
for t in 320 352 384; do python bench.py --workload div_p4 --steps 10 --warmup 3 --no-e2e --no-cpu --param threads=$t > gpurun_out/c2_div_$t.json 2> gpurun_out/c2_div_$t.err; done
