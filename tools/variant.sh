#!/bin/bash
# A/B builds of kernel variants behind compile-time switches (profiles/r02_ab_dmma.md, "parameter sweeps"):
#   tools/variant.sh build NAME "-DFNSM_KLP=5 ..."   -> tools/_ab/libv_NAME.so (opmat.cu recompiled with the flags,
#                                                       linked with the in-tree objects of the other sources)
#   tools/variant.sh run "workload ..." NAME ...      -> bench.py --no-suite on every variant, twice, interleaved
# Switches that exist in the sources: FNSM_KLP (left-over partial sums of the divergence kernel, default 6),
# FNSM_LIFT_LP (the same for the lift kernel, default 2).  tools/_ab/ is scratch (built libraries travel with gpurun).
cd "$(dirname "$0")/.." || exit 1
OBJ=feinsum_b200/csrc/_obj
mkdir -p tools/_ab
case "$1" in
  build)
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
         -I include $3 -c feinsum_b200/csrc/opmat.cu -o tools/_ab/opmat_$2.o || exit 1
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/_ab/libv_$2.so tools/_ab/opmat_$2.o \
         $OBJ/abi.*.o $OBJ/generic.*.o $OBJ/hex_deriv.*.o $OBJ/peaks.*.o $OBJ/tensor_product.*.o -lcudart
    rm -f tools/_ab/opmat_$2.o; ls -la tools/_ab/libv_$2.so ;;
  run)
    wl=$2; shift 2
    for rep in 1 2; do for v in "$@"; do for w in $wl; do
      FNSM_B200_LIB=$PWD/tools/_ab/libv_$v.so timeout 100 python bench.py --workload $w --no-e2e --no-cpu --no-suite --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w', '$v', round(l['ms_per_step'], 4), round(l['roofline']['roofline_frac'], 4))"
    done; done; done ;;
  *) echo "usage: $0 build NAME FLAGS | run 'workloads' NAME..."; exit 1 ;;
esac
