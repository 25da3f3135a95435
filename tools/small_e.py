#!/usr/bin/env python
"""Where does the time go at small E (BASELINE config 1: E = 100 000)?

For the fp64 p = 4 kernels: time per launch against the number of elements, (a) enqueued from Python back to back on a
stream and (b) replayed from a CUDA graph (no host in the loop), plus the host's own enqueue cost.  A straight-line fit
T(E) = T0 + E / rate separates the fixed cost of a launch (prologue, table staging, tail) from the streaming rate.

    python tools/small_e.py [--kinds grad div lift] [--param K=V ...] > profiles/rNN_small_e.md
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import feinsum_b200 as f  # noqa: E402
from feinsum_b200.codegen import generate_cuda  # noqa: E402
from tests import einsums as E  # noqa: E402

FLOPS = {"grad": 7980.0, "div": 7980.0, "lift": 17040.0}
BYTES = {"grad": 1192.0, "div": 1192.0, "lift": 3072.0}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--kinds", nargs="+", default=["grad", "div", "lift"])
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--sizes", type=int, nargs="+",
                    default=[2368, 23680, 47360, 71040, 94720, 100000, 118400, 189440, 400000])
    ap.add_argument("--param", action="append", default=[], metavar="K=V")
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    params = {k: int(v) for k, v in (p.split("=") for p in args.param)}
    side = torch.cuda.Stream()
    cq = f.CudaQueue(0, stream=side)
    builders = {"grad": E.grad, "div": E.div, "lift": E.lift_fe}
    peak = 37.1e12 if args.dtype == "float64" else 73.5e12
    tdt = getattr(torch, args.dtype)
    print(f"# small-E decomposition ({args.dtype}, params {params or 'default'}), {torch.cuda.get_device_name(0)}\n")
    for kind in args.kinds:
        e = builders[kind](dtype=args.dtype)
        prog = generate_cuda(e)
        if params:
            prog = prog.with_params(**params)
        ex = prog.executor(cq)
        print(f"## {kind}\n")
        print("| E | stream loop us/launch | graph us/launch | roofline us | graph frac | host us/call |")
        print("|---|---|---|---|---|---|")
        xs, ys = [], []
        for n in args.sizes:
            arrs = {k: torch.rand(tuple(n if not isinstance(d, int) else d for d in s), dtype=tdt, device="cuda")
                    for k, s in e.arg_to_shape.items()}
            torch.cuda.synchronize()
            with torch.cuda.stream(side):
                evt, outs = ex(cq, **arrs)
                evt.wait()
                full = dict(arrs)
                full.update(outs)
                for _ in range(20):
                    ex(cq, **full)
                side.synchronize()
                # host cost: enqueue without waiting
                t0 = time.perf_counter()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(side)
                for _ in range(args.reps):
                    ex(cq, **full)
                b.record(side)
                t1 = time.perf_counter()
                b.synchronize()
                loop_us = a.elapsed_time(b) * 1e3 / args.reps
                host_us = (t1 - t0) * 1e6 / args.reps
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(args.reps):
                    ex(cq, **full)
            with torch.cuda.stream(side):
                g.replay()
                side.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(side)
                g.replay()
                b.record(side)
                b.synchronize()
                graph_us = a.elapsed_time(b) * 1e3 / args.reps
            roof = max(FLOPS[kind] * n / peak, BYTES[kind] * (0.5 if args.dtype == "float32" else 1.0) * n / 6561.6e9) * 1e6
            print(f"| {n} | {loop_us:.2f} | {graph_us:.2f} | {roof:.2f} | {roof / graph_us:.3f} | {host_us:.1f} |")
            xs.append(n)
            ys.append(graph_us)
            del g
        A = np.vstack([np.ones(len(xs)), np.array(xs, float)]).T
        (t0_, slope), *_ = np.linalg.lstsq(A, np.array(ys), rcond=None)
        print(f"\nfit (graph): T(E) = {t0_:.2f} us + E x {slope * 1e3:.4f} ns  "
              f"(streaming rate = {FLOPS[kind] / slope * 1e-6:.1f} TFLOP/s)\n")


if __name__ == "__main__":
    main()
