#!/usr/bin/env python
"""Roofline table of every BASELINE einsum through the reference-style API
(``feinsum_b200.measure``): GFLOP/s, GB/s and % of the B200 roofline at E = 100 000
(BASELINE config 1 size; the working set fits the 126 MB L2) and E = 4 000 000 (HBM).

    python tools/report.py [--sizes 100000 4000000] > profiles/rNN_report.md
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import feinsum_b200 as f  # noqa: E402
from feinsum_b200 import _cabi, measure  # noqa: E402
from feinsum_b200.data import device_info  # noqa: E402
from tests import einsums as E  # noqa: E402

IDENTITY = lambda t_unit, insn_match=None, kernel_name=None: t_unit  # noqa: E731


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[100_000, 4_000_000])
    ap.add_argument("--orders", action="store_true",
                    help="tets of order p = 1..3 (4/10/20 volume dofs, 3/6/10 face dofs) instead of the BASELINE shapes")
    ap.add_argument("--secs", type=float, default=0.5, help="minimum timed seconds per row")
    args = ap.parse_args()
    cq = f.CudaQueue(0)
    fp64 = max(_cabi.measure_peak(0), _cabi.measure_peak(3))
    fp32 = _cabi.measure_peak(1)
    # the same micro-kernels after ~0.7 s of load: the board sits at its power cap, which is where timeit's >= 0.5 s
    # loops run (the burst figure stays the denominator of the "% of roofline" column)
    sus = {"float64": min(fp64, max(_cabi.measure_peak(16), _cabi.measure_peak(19))), "float32": min(fp32, _cabi.measure_peak(17))}
    device_info.register_measured_peaks(cq.device.name, float64=fp64, float32=fp32)
    bw = device_info.DEV_TO_PEAK_BW[cq.device.name]
    measure.N_MIN_SIM_SECS = args.secs
    print(f"# Roofline report, {cq.device.name}: FP64 {fp64:.0f} GFLOP/s, FP32 {fp32:.0f} GFLOP/s (measured in this run), "
          f"HBM {bw:.1f} GB/s (MEASURED_PEAKS.json)\n")
    print("`feinsum_b200.measure.timeit` (validation gate, 5 warm-ups, CUDA events); FLOPs = flop-optimal contraction "
          "path, bytes = every operand and output once.\n")
    print(f"Sustained FP peaks (after 0.7 s of load): FP64 {sus['float64']:.0f}, FP32 {sus['float32']:.0f} GFLOP/s -> last column.\n")
    print("| einsum | dtype | E | ms | GFLOP/s | GB/s | roofline GFLOP/s | % of roofline | % of the roofline at the sustained FP peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    cases = [("grad xre,rij,ej->xei", E.grad), ("div xre,rij,xej->ei", E.div),
             ("lift ifj,fe,fej->ei b=4", E.lift_fe), ("lift ef,fij,fej->ei b=4", E.lift_ef)]
    rows = []
    for n in args.sizes:
        if args.orders:
            for p, nd, nfd in ((1, 4, 3), (2, 10, 6), (3, 20, 10)):
                for dt in ("float64", "float32"):
                    rows.append((f"grad p={p}", E.grad(dtype=dt, ndof=nd), dt, n))
                    rows.append((f"div p={p}", E.div(dtype=dt, ndof=nd), dt, n))
                    rows.append((f"lift ifj,fe,fej->ei b=4 p={p}", E.lift_fe(dtype=dt, nvol=nd, nfd=nfd), dt, n))
            continue
        for name, builder in cases:
            for dt in ("float64", "float32"):
                rows.append((name, builder(dtype=dt), dt, n))
        for dt in ("float64", "float32"):
            rows.append(("tensor-product eabc,ia->eibc p=7", E.tensor_product(0, 8, dt), dt, n))
    for name, e, dt, n in rows:
        t = measure.timeit(e, transform=IDENTITY, cq=cq, long_dim_length=n)
        gops = sum(measure._get_giga_ops_from_einsum(e, n).values())
        gb = measure._get_footprint_gbytes(e, n)
        roof = measure.get_roofline_flop_rate(e, cq.device.name, n)[np.dtype(dt)]
        roof_sus = min(sus[dt], gops / gb * bw)       # the roofline with its FP leg at the sustained peak
        print(f"| {name} | {dt} | {n} | {t * 1e3:.4f} | {gops / t:.0f} | {gb / t:.0f} | {roof:.0f} | {100 * gops / t / roof:.1f} | "
              f"{100 * gops / t / roof_sus:.1f} |", flush=True)


if __name__ == "__main__":
    main()
