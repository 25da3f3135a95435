#!/usr/bin/env python
"""
bench.py -- BASELINE.json metric: GFLOP/s & HBM GB/s (% of B200 roofline) per
DG einsum, beside the CPU restatement of the reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One *step* = one execution of the einsum over the whole element batch
(BASELINE config 2 by default: DG divergence ``xre,rij,xej->ei``, p = 4 tets,
fp64, E = 4 000 000 elements per GPU; 4.77 GB of operands >> 126 MB L2, so
every step streams from HBM -- no L2 flush needed).  N > 1: launched under
torchrun, one rank per GPU, element axis sharded (every rank owns E elements:
weak scaling), no collective on the data path; time = max over ranks.

Printed JSON (one line, rank 0):  value = whole-job GFLOP/s with operands
resident in HBM; ``e2e`` = same metric through the host-buffer API
(``HostExecutor``: pinned numpy in, numpy out, H2D/D2H inside the timed
region); ``roofline`` = the kernel against the measured FP64 peak
(``fnsm_b200_measure_peak`` in this run -- MEASURED_PEAKS.json only carries
HBM and bf16) and against the measured HBM copy bandwidth; ``cpu_baseline`` =
the oracle's C/OpenMP restatement of the reference's generated loop nest on the
host cores (loopy -> pocl cannot run in this image).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "GFLOP/s per DG einsum (opt_einsum-path flops), with HBM GB/s and % of B200 roofline"

WORKLOADS = {
    # name: (builder name in this file, default E per GPU, description)
    "div_p4": ("DG divergence xre,rij,xej->ei p=4 tets fp64", 4_000_000),
    "grad_p4": ("DG gradient xre,rij,ej->xei p=4 tets fp64", 4_000_000),
    "lift_p4": ("DG face-mass lift ifj,fe,fej->ei b=4 p=4 tets fp64", 4_000_000),
    "tp_p7": ("tensor-product eabc,ia->eibc p=7 hexes fp64", 4_000_000),
    "div_p4_f32": ("DG divergence xre,rij,xej->ei p=4 tets fp32", 4_000_000),
    "grad_p4_f32": ("DG gradient xre,rij,ej->xei p=4 tets fp32", 4_000_000),
    "lift_p4_f32": ("DG face-mass lift ifj,fe,fej->ei b=4 p=4 tets fp32", 4_000_000),
    "wave_p4": ("wave_3d_p4 operator: div(v) + grad(u) + 4-field lift in one call, fp64", 4_000_000),
    "wave_p4_f32": ("wave_3d_p4 operator: div(v) + grad(u) + 4-field lift in one call, fp32", 4_000_000),
}
# lower-order tets: p = 1..3 -> (volume dofs, face dofs)
ORDERS = {1: (4, 3), 2: (10, 6), 3: (20, 10), 4: (35, 15)}
for _p in (1, 2, 3):
    for _k, _sub in (("div", "xre,rij,xej->ei"), ("grad", "xre,rij,ej->xei"), ("lift", "ifj,fe,fej->ei b=4")):
        WORKLOADS[f"{_k}_p{_p}"] = (f"DG {_k} {_sub} p={_p} tets fp64", 4_000_000)
        WORKLOADS[f"{_k}_p{_p}_f32"] = (f"DG {_k} {_sub} p={_p} tets fp32", 4_000_000)


def build_einsum(name: str):
    import feinsum_b200 as f

    dt = "float32" if name.endswith("_f32") else "float64"
    base = name.replace("_f32", "")
    kind, _, order = base.partition("_p")
    if kind in ("div", "grad", "lift") and order in ("1", "2", "3", "4"):
        nd, nfd = ORDERS[int(order)]
        if kind == "div":
            return f.einsum("xre,rij,xej->ei", f.array("J", (3, 3, "E"), dt),
                            f.array("D", (3, nd, nd), dt), f.array("u", (3, "E", nd), dt))
        if kind == "grad":
            return f.einsum("xre,rij,ej->xei", f.array("J", (3, 3, "E"), dt),
                            f.array("D", (3, nd, nd), dt), f.array("u", ("E", nd), dt))
        return f.batched_einsum(
            "ifj,fe,fej->ei",
            [[f.array("L", (nd, 4, nfd), dt), f.array("Jface", (4, "E"), dt),
              f.array(f"F_{k}", (4, "E", nfd), dt)] for k in range(4)])
    if base == "tp_p7":
        return f.einsum("eabc,ia->eibc", f.array("A", ("E", 8, 8, 8), dt), f.array("M", (8, 8), dt))
    if base == "wave_p4":
        return None          # three einsums behind one call: see feinsum_b200/wave3d.py
    raise SystemExit(f"unknown workload {name}")


def concrete(shape, n):
    return tuple(int(d) if isinstance(d, (int, np.integer)) else n for d in shape)


# ------------------------------------------------------------------ clocks --
class ClockSampler:
    """Samples SM clock and throttle reasons via NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# --------------------------------------------------------------- cpu legs ---
def cpu_reference_leg(einsum, flops_per_elem: float, sample_e: int, steps: int, warmup: int):
    """Times the oracle's C/OpenMP loop nest (trivial schedule, as generate_loopy +
    identity transform emits it) on all host cores.  Returns (GFLOP/s, cores, ms/step)."""
    from oracle import cgen, np_oracle

    cores = len(os.sched_getaffinity(0))
    # torchrun exports OMP_NUM_THREADS=1; this leg is meant to use every host core
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)   # if the runtime was initialised already
    except OSError:
        pass
    kern = cgen.CKernel(einsum)
    ins = np_oracle.generate_input_arrays(einsum, sample_e, 0)
    outs = [np.empty(s, dtype=np.result_type(*[a.dtype for a in row]))
            for s, row in zip(kern.out_shapes(sample_e), einsum.args)]
    for _ in range(max(1, warmup)):
        kern(sample_e, ins, outs)
    t0 = time.perf_counter()
    for _ in range(steps):
        kern(sample_e, ins, outs)
    dt = (time.perf_counter() - t0) / steps
    return flops_per_elem * sample_e / dt * 1e-9, cores, dt * 1e3


def cpu_reference_leg_wave(dtype: str, sample_e: int, steps: int, warmup: int):
    """CPU port of the wave operator = its three loop nests run one after the other."""
    from feinsum_b200 import measure, wave3d

    secs = 0.0
    cores = 1
    for e in wave3d.wave3d_einsums("float32" if dtype == "f32" else "float64").values():
        fl = sum(measure.get_flops_per_dtype(e, 1_000_000).values()) / 1e6
        gf, cores, _ = cpu_reference_leg(e, fl, sample_e, steps, warmup)
        secs += fl * sample_e / (gf * 1e9)
    return wave3d.FLOPS_PER_ELEMENT * sample_e / secs * 1e-9, cores, secs * 1e3


def profiled_traffic(workload: str):
    """DRAM bytes of one launch of the workload's kernel from the committed ncu capture
    (profiles/rNN_ncu_<workload>.txt: dram__bytes_read.sum + dram__bytes_write.sum); None if absent."""
    import glob
    import re

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_ncu_{workload}.txt")))
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot = 0.0
    text = open(files[-1]).read()
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(rf"{re.escape(key)} = ([0-9.]+) (\w+)", text)
        if not m:
            return None, None
        tot += float(m.group(1)) * unit.get(m.group(2), 1.0)
    return tot, os.path.relpath(files[-1], ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="div_p4", choices=sorted(WORKLOADS))
    ap.add_argument("--elements", type=int, default=0, help="elements per GPU (default: workload's)")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 dmma, 2 simt")
    ap.add_argument("--param", action="append", default=[], metavar="K=V",
                    help="launch parameter of the kernel (threads=384, ...); may repeat")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from feinsum_b200 import measure

    einsum = build_einsum(args.workload)
    is_wave = einsum is None
    descr, default_e = WORKLOADS[args.workload]
    E = args.elements or default_e
    dtype = "f32" if args.workload.endswith("_f32") else "f64"
    if is_wave:
        from feinsum_b200 import wave3d

        flops_per_elem = float(wave3d.FLOPS_PER_ELEMENT)
        bytes_per_elem = float(wave3d.BYTES_PER_ELEMENT[np.dtype("float32" if dtype == "f32" else "float64")])
    else:
        flops_per_elem = sum(measure.get_flops_per_dtype(einsum, 1_000_000).values()) / 1e6
        bytes_per_elem = (measure.get_footprint_bytes(einsum, 2_000_000)
                          - measure.get_footprint_bytes(einsum, 1_000_000)) / 1e6
    config = {
        "workload": f"{descr}, {E} elements per GPU (BASELINE configs[1] family)",
        "elements_per_gpu": E,
        "flops_per_element": flops_per_elem,
        "bytes_per_element": bytes_per_elem,
        "l2_policy": "operands >> L2 (4.77 GB vs 126 MB): inputs larger than L2, no flush",
        "parallelism": f"element axis sharded over {max(world, args.gpus)} GPU(s), no collective",
    }

    # ------------------------------------------------------ reference arm ---
    if args.impl == "reference":
        if rank != 0:
            return
        sample_e = 400_000
        if is_wave:
            gf, cores, ms = cpu_reference_leg_wave(dtype, sample_e, args.steps, args.warmup)
        else:
            gf, cores, ms = cpu_reference_leg(einsum, flops_per_elem, sample_e, args.steps, args.warmup)
        line = {
            "impl": "reference", "metric": METRIC,
            "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": config,
            "cpu_baseline": {
                "value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                "sample": f"{sample_e} elements per step: C/OpenMP restatement of the loop nest "
                          "generate_loopy emits (trivial schedule, -O3 -ffast-math -fopenmp); "
                          "the reference's loopy->pocl path cannot run in this image",
            },
            "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------ B200 arm --
    import torch

    import feinsum_b200 as f
    from feinsum_b200 import _cabi
    from feinsum_b200.codegen import generate_cuda
    from feinsum_b200.data import device_info
    from feinsum_b200.host_exec import HostExecutor, pinned_empty
    from feinsum_b200 import wave3d

    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    cq = f.CudaQueue(local_rank)
    dev = cq.torch_device

    params = {k: int(v) for k, v in (kv.split("=") for kv in args.param)}
    if args.variant > 0:
        params["variant"] = args.variant
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    tdt = torch.float32 if dtype == "f32" else torch.float64
    if is_wave:
        prog = None
        ex = wave3d.Wave3DExecutor(cq, "float32" if dtype == "f32" else "float64", **params)
        in_shapes, out_shapes = wave3d.shapes(E)
        kernel_id = "wave3d"
    else:
        prog = generate_cuda(einsum)
        if params:
            prog = prog.with_params(**params)
        ex = prog.executor(cq)
        in_shapes = {n: concrete(s, E) for n, s in sorted(einsum.arg_to_shape.items())}
        out_shapes = {n: concrete(einsum.shape, E) for n in einsum.output_names}
        kernel_id = prog.kernel_id
    arrays = {n: torch.rand(s, dtype=tdt, device=dev, generator=gen) for n, s in sorted(in_shapes.items())}
    outs = {n: torch.zeros(s, dtype=tdt, device=dev) for n, s in out_shapes.items()}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # measured peaks for the roofline (this board, this run): burst figure for a short timed region,
    # sustained (power-capped) figure when the timed region is long enough to pull the power cap
    peak_burst = _cabi.measure_peak(3 if dtype == "f64" else 1)   # DMMA fp64 / FFMA2 fp32
    if dtype == "f64":
        peak_burst = max(peak_burst, _cabi.measure_peak(0))
    peak_fp = peak_burst
    peak_kind = "burst"
    hbm_peak = device_info.DEV_TO_PEAK_BW.get("NVIDIA B200", 6561.6)
    hbm_src = "MEASURED_PEAKS.json" if device_info._hbm is not None else "fallback table"

    for _ in range(args.warmup):
        ex(cq, **arrays, **outs)
    barrier()
    launches0 = _cabi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        start.record(cq.torch_stream)
        for _ in range(args.steps):
            ex(cq, **arrays, **outs)
        stop.record(cq.torch_stream)
        stop.synchronize()
        barrier()
    launches = _cabi.launch_count() - launches0
    ms_total = start.elapsed_time(stop)
    if ms_total > 250.0 or "sw_power_cap" in clocks.summary()["reasons"]:
        peak_fp = min(peak_burst, _cabi.measure_peak((3 if dtype == "f64" else 1) + 16))
        peak_kind = "sustained (power-capped, measured after 0.7 s of load)"
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = flops_per_elem * E * world / (ms_step * 1e-3) * 1e-9
    gbs = bytes_per_elem * E * world / (ms_step * 1e-3) * 1e-9

    # per-kernel roofline (this rank's kernel)
    my_ms = ms_total / args.steps
    launches_per_step = max(1, launches // args.steps)
    ach_tflops = flops_per_elem * E / (my_ms * 1e-3) * 1e-12
    ach_gbs = bytes_per_elem * E / (my_ms * 1e-3) * 1e-9
    t_flop = flops_per_elem * E / (peak_fp * 1e9)
    t_mem = bytes_per_elem * E / (hbm_peak * 1e9)
    if t_flop >= t_mem:
        roof = {"bound": "tensor", "achieved": ach_tflops, "peak": peak_fp * 1e-3,
                "unit": "TFLOP/s", "frac": ach_tflops / (peak_fp * 1e-3)}
    else:
        roof = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach_gbs / hbm_peak}
    traffic, traffic_src = (profiled_traffic(args.workload) if E == default_e and not params else (None, None))
    roof.update({
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_kind": peak_kind, "peak_burst": peak_burst * 1e-3,
        "peak_source": (f"{'FP64 DMMA/DFMA' if dtype == 'f64' else 'FP32 FFMA2'} {peak_kind} peak measured in this run by "
                        f"fnsm_b200_measure_peak (MEASURED_PEAKS.json has no {dtype} figure); "
                        f"HBM {hbm_peak} GB/s from {hbm_src}"),
        "t_roof_ms": max(t_flop, t_mem) * 1e3, "roofline_frac": max(t_flop, t_mem) / (my_ms * 1e-3),
        "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "frac": ach_gbs / hbm_peak},
        "fp": {"achieved": ach_tflops, "peak": peak_fp * 1e-3, "frac": ach_tflops / (peak_fp * 1e-3)},
        "launches_per_step": launches_per_step,
        "algorithmic_bytes_per_launch": bytes_per_elem * E / launches_per_step,
        "algorithmic_flops_per_launch": flops_per_elem * E / launches_per_step,
    })

    # ---------------------------------------------------------------- e2e ---
    e2e = None
    if not args.no_e2e and is_wave:
        # host buffers in, host buffers out, one stream: H2D of all ten operands, the fused
        # call, D2H of the six results (no chunk pipelining for the three-einsum operator yet)
        npdt = np.float32 if dtype == "f32" else np.float64
        host_in = {n: pinned_empty(s, npdt) for n, s in in_shapes.items()}
        for n in host_in:
            torch.from_numpy(host_in[n]).copy_(arrays[n])
        host_out = {n: pinned_empty(s, npdt) for n, s in out_shapes.items()}
        dev_in = {n: torch.empty_like(a) for n, a in arrays.items()}

        def e2e_step():
            with torch.cuda.stream(cq.torch_stream):
                for n in dev_in:
                    dev_in[n].copy_(torch.from_numpy(host_in[n]), non_blocking=True)
                ex(cq, **dev_in, **outs)
                for n in host_out:
                    torch.from_numpy(host_out[n]).copy_(outs[n], non_blocking=True)
            cq.finish()

        e2e_steps = 3
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": flops_per_elem * E * world / dt * 1e-9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": int(sum(a.nbytes for a in host_in.values())),
               "d2h_bytes_per_step": int(sum(a.nbytes for a in host_out.values())),
               "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "note": "pinned numpy in/out; H2D, fused call, D2H in order on one stream"}
    elif not args.no_e2e:
        host_in = {}
        for name, shape in sorted(einsum.arg_to_shape.items()):
            h = pinned_empty(concrete(shape, E), np.float32 if dtype == "f32" else np.float64)
            torch.from_numpy(h).copy_(arrays[name])
            host_in[name] = h
        host_out = {n: pinned_empty(concrete(einsum.shape, E),
                                    np.float32 if dtype == "f32" else np.float64)
                    for n in einsum.output_names}
        hx = HostExecutor(prog, cq)
        e2e_steps = max(3, min(args.steps, 5))
        hx(outputs=host_out, **host_in)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hx(outputs=host_out, **host_in)
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": flops_per_elem * E * world / dt * 1e-9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": hx.h2d_bytes, "d2h_bytes_per_step": hx.d2h_bytes,
               "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "note": "HostExecutor: pinned numpy in/out, chunked H2D | kernel | D2H on 3 streams"}
        # spot parity of the host path against the device path
        ref = outs[einsum.output_names[0]][..., :1].cpu().numpy()
        got = host_out[einsum.output_names[0]][..., :1]
        if not np.allclose(got, ref, rtol=1e-12 if dtype == "f64" else 1e-5):
            raise SystemExit("host path and device path disagree")

    # --------------------------------------------------------- cpu baseline -
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_e = 400_000
        if is_wave:
            gf, cores, ms = cpu_reference_leg_wave(dtype, sample_e, 3, 1)
        else:
            gf, cores, ms = cpu_reference_leg(einsum, flops_per_elem, sample_e, 3, 1)
        cpu = {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port",
               "sample": f"{sample_e} elements x 3 steps of the C/OpenMP restatement of "
                         f"generate_loopy's loop nest ({ms:.1f} ms/step)"}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": config, "gbs": gbs, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "kernel": kernel_id, "device": cq.device.name,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
