#!/usr/bin/env python
"""
bench.py -- BASELINE.json metric: GFLOP/s & HBM GB/s (% of B200 roofline) per
DG einsum at 1/2/4/8 GPUs, beside the CPU restatement of the reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME]
                    [--impl reference] [--scaling weak|strong] [--no-suite]

Headline (``value``, ``roofline``, ``e2e``, ``cpu_baseline``): one *step* = one
execution of ``--workload`` over the whole element batch (default BASELINE
config 2: DG divergence ``xre,rij,xej->ei``, p = 4 tets, fp64, E = 4 000 000
elements per GPU; 4.77 GB of operands >> 126 MB L2, so every step streams from
HBM).  N > 1: launched under torchrun, one rank per GPU, element axis sharded,
no collective on the data path; time = max over ranks.

``workloads``: unless ``--no-suite``, the same line carries every other
BASELINE config timed the same way (fewer steps): grad p4 at E = 100 000
(config 1; L2-resident and with L2 flushed between launches) and 4 M, div / lift
at 100 000, lift p4 (config 3), the wave_3d_p4 operator in fp64 and fp32
(config 4), tensor-product p7 (config 5: 4 M elements per GPU = 32 M at N = 8)
and the three fp32 DG kernels -- each with ms/step, GFLOP/s, GB/s and its
roofline fraction against the burst and the sustained peaks.
``strong``: div p4, wave_3d_p4 and tensor-product at a FIXED total of
16 000 000 elements split over the N ranks.

``e2e`` = the headline metric through the host-buffer API (``HostExecutor``:
pinned numpy in, numpy out, H2D/D2H inside the timed region), with the raw
pinned-copy time of the same bytes on the same ranks as ``pcie_floor_ms``.
``cpu_baseline`` / ``--impl reference`` = the oracle's C/OpenMP restatement of
the reference's generated loop nest on the host cores (loopy -> pocl cannot run
in this image); the GPU arm runs the reference arm's code in a subprocess so
both legs follow one policy.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "GFLOP/s per DG einsum (opt_einsum-path flops), with HBM GB/s and % of B200 roofline"
L2_BYTES = 126e6
STRONG_TOTAL = 16_000_000

WORKLOADS = {
    # name: (description, default elements per GPU, BASELINE.json config it belongs to)
    "div_p4": ("DG divergence xre,rij,xej->ei p=4 tets fp64", 4_000_000, "configs[1]"),
    "grad_p4": ("DG gradient xre,rij,ej->xei p=4 tets fp64", 4_000_000, "configs[0] einsum at the configs[1] size"),
    "lift_p4": ("DG face-mass lift ifj,fe,fej->ei b=4 p=4 tets fp64", 4_000_000, "configs[2]"),
    "tp_p7": ("tensor-product eabc,ia->eibc p=7 hexes fp64", 4_000_000, "configs[4] (4 M elements per GPU)"),
    "div_p4_f32": ("DG divergence xre,rij,xej->ei p=4 tets fp32", 4_000_000, "configs[1] einsum in fp32"),
    "grad_p4_f32": ("DG gradient xre,rij,ej->xei p=4 tets fp32", 4_000_000, "configs[0] einsum in fp32"),
    "lift_p4_f32": ("DG face-mass lift ifj,fe,fej->ei b=4 p=4 tets fp32", 4_000_000, "configs[2] einsum in fp32"),
    "wave_p4": ("wave_3d_p4 operator: div(v) + grad(u) + 4-field lift in one call, fp64", 4_000_000, "configs[3] fp64"),
    "wave_p4_f32": ("wave_3d_p4 operator: div(v) + grad(u) + 4-field lift in one call, fp32", 4_000_000, "configs[3] fp32"),
    "se_p4": ("div components se,sij,ej->ei: 3 rows J_b(3,E) R(3,35,35) u_b(E,35) p=4 tets fp64 (shared operator)",
              4_000_000, "SURVEY 8(f3); reference test/test_codegen.py:34-60"),
    "se_face_p4": ("face mass se,sij,ej->ei: 4 rows J(4,E) R(4,15,15) v_k(E,15) fp64 (shared operator)",
                   4_000_000, "SURVEY 8(f3); reference test/test_codegen.py:63-88"),
    "hexd_p7": ("fused hex derivative eabc,ia->eibc + eabc,ib->eaic + eabc,ic->eabi p=7 fp64 (A read once)",
                4_000_000, "SURVEY 8(f3)"),
}
# lower-order tets: p = 1..3 -> (volume dofs, face dofs)
ORDERS = {1: (4, 3), 2: (10, 6), 3: (20, 10), 4: (35, 15)}
for _p in (1, 2, 3):
    for _k, _sub in (("div", "xre,rij,xej->ei"), ("grad", "xre,rij,ej->xei"), ("lift", "ifj,fe,fej->ei b=4")):
        WORKLOADS[f"{_k}_p{_p}"] = (f"DG {_k} {_sub} p={_p} tets fp64", 4_000_000, "SURVEY 8(f3)")
        WORKLOADS[f"{_k}_p{_p}_f32"] = (f"DG {_k} {_sub} p={_p} tets fp32", 4_000_000, "SURVEY 8(f3)")

# what the default command times besides the headline: (workload, elements per GPU or None = default, flush L2?)
SUITE = [
    ("grad_p4", 100_000, False), ("grad_p4", 100_000, True), ("div_p4", 100_000, False),
    ("lift_p4", 100_000, False),
    ("grad_p4", None, False), ("lift_p4", None, False), ("wave_p4", None, False),
    ("wave_p4_f32", None, False), ("tp_p7", None, False),
    ("grad_p4_f32", None, False), ("div_p4_f32", None, False), ("lift_p4_f32", None, False),
    # shared-operator families on the tensor path; the alignment cliff (odd E / E % 4 != 0)
    ("se_p4", None, False), ("se_face_p4", None, False),
    ("div_p4", 4_000_001, False), ("lift_p4", 4_000_001, False),
    ("div_p4_f32", 4_000_001, False), ("lift_p4_f32", 4_000_002, False),
]
STRONG_SUITE = ["div_p4", "wave_p4", "tp_p7"]


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
    if base == "se_p4":
        return f.batched_einsum("se,sij,ej->ei", [[f.array(f"J{c}", (3, "E"), dt), f.array("R", (3, 35, 35), dt),
                                                   f.array(f"u{c}", ("E", 35), dt)] for c in "xyz"])
    if base == "se_face_p4":
        return f.batched_einsum("se,sij,ej->ei", [[f.array("J", (4, "E"), dt), f.array("R", (4, 15, 15), dt),
                                                   f.array(f"v{k}", ("E", 15), dt)] for k in range(4)])
    if base == "tp_p7":
        return f.einsum("eabc,ia->eibc", f.array("A", ("E", 8, 8, 8), dt), f.array("M", (8, 8), dt))
    if base in ("wave_p4", "hexd_p7"):
        return None          # several einsums behind one call: feinsum_b200/wave3d.py, hexderiv.py
    raise SystemExit(f"unknown workload {name}")


def concrete(shape, n):
    return tuple(int(d) if isinstance(d, (int, np.integer)) else n for d in shape)


def dtype_of(name: str) -> str:
    return "f32" if name.endswith("_f32") else "f64"


class Work:
    """Per-element work model + program of one workload (no GPU needed to construct)."""

    def __init__(self, name: str, params: dict | None = None):
        from feinsum_b200 import measure

        self.name = name
        self.descr, self.default_e, self.baseline_cfg = WORKLOADS[name]
        self.dtype = dtype_of(name)
        self.np_dtype = np.float32 if self.dtype == "f32" else np.float64
        self.params = dict(params or {})
        self.einsum = build_einsum(name)
        base = name.replace("_f32", "")
        if base == "wave_p4":
            from feinsum_b200 import wave3d

            self.flops = float(wave3d.FLOPS_PER_ELEMENT)
            self.bytes = float(wave3d.BYTES_PER_ELEMENT[np.dtype(self.np_dtype)])
            self.program = wave3d.Wave3DProgram(self.np_dtype, **self.params)
            self.cpu_einsums = list(wave3d.wave3d_einsums(self.np_dtype).values())
        elif base == "hexd_p7":
            from feinsum_b200 import hexderiv

            self.flops = float(hexderiv.FLOPS_PER_ELEMENT)
            self.bytes = float(hexderiv.BYTES_PER_ELEMENT[np.dtype(self.np_dtype)])
            self.program = hexderiv.HexDerivProgram(self.np_dtype, **self.params)
            self.cpu_einsums = list(hexderiv.hexderiv_einsums(self.np_dtype).values())
        else:
            from feinsum_b200.codegen import generate_cuda

            self.flops = sum(measure.get_flops_per_dtype(self.einsum, 1_000_000).values()) / 1e6
            self.bytes = (measure.get_footprint_bytes(self.einsum, 2_000_000)
                          - measure.get_footprint_bytes(self.einsum, 1_000_000)) / 1e6
            self.program = generate_cuda(self.einsum)
            if self.params:
                self.program = self.program.with_params(**self.params)
            self.cpu_einsums = [self.einsum]

    def shapes(self, E: int):
        if self.einsum is None:
            spec = self.program.host_spec()
            return ({n: concrete(s, E) for n, s in sorted(spec.in_shapes.items())},
                    {n: concrete(s, E) for n, s in spec.out_shapes.items()})
        es = self.einsum
        return ({n: concrete(s, E) for n, s in sorted(es.arg_to_shape.items())},
                {n: concrete(es.shape, E) for n in es.output_names})

    def l2_policy(self, E: int, flush: bool) -> str:
        ws = self.bytes * E
        if flush:
            return (f"working set {ws / 1e6:.0f} MB <= L2: L2 flushed between timed launches "
                    "(256 MB buffer rewritten; each launch timed by its own event pair)")
        if ws > 4 * L2_BYTES:
            return f"operands >> L2 ({ws / 1e9:.2f} GB vs 126 MB): inputs larger than L2, no flush"
        if ws > L2_BYTES:
            return f"working set {ws / 1e6:.0f} MB vs 126 MB L2: partially cached, launched back to back WITHOUT flush"
        return (f"working set {ws / 1e6:.0f} MB <= 126 MB L2: launched back to back WITHOUT flush "
                "(L2-resident, the regime of the reference's database size E = 100 000)")


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
def cpu_reference_leg(einsums, flops_per_elem: float, sample_e: int, steps: int, warmup: int):
    """Times the oracle's C/OpenMP loop nests (trivial schedule, as generate_loopy +
    identity transform emits them; one nest per einsum of the workload, run one after the
    other) on all host cores.  Returns (GFLOP/s, cores, ms/step)."""
    from oracle import cgen, np_oracle

    cores = len(os.sched_getaffinity(0))
    # torchrun exports OMP_NUM_THREADS=1; this leg is meant to use every host core
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)   # if the runtime was initialised already
    except OSError:
        pass
    legs = []
    for es in einsums:
        kern = cgen.CKernel(es)
        ins = np_oracle.generate_input_arrays(es, sample_e, 0)
        outs = [np.empty(s, dtype=np.result_type(*[a.dtype for a in row]))
                for s, row in zip(kern.out_shapes(sample_e), es.args)]
        legs.append((kern, ins, outs))

    def step():
        for kern, ins, outs in legs:
            kern(sample_e, ins, outs)

    for _ in range(max(1, warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return flops_per_elem * sample_e / dt * 1e-9, cores, dt * 1e3


def reference_sample_elements(work: Work, E: int) -> int:
    """Elements per step of the CPU arm: the labelled E when inputs + outputs fit comfortably in
    host RAM and a step stays within a few seconds, else a bounded sample (stated in the line)."""
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 8e9
    n = E
    while n > 100_000 and (work.bytes * n * 2.5 > avail or work.flops * n > 4e10):
        n //= 2
    return n


def reference_arm(work: Work, E: int, n_gpus: int, steps: int, warmup: int, sample_e: int | None) -> dict:
    sample_e = sample_e or reference_sample_elements(work, E)
    gf, cores, ms = cpu_reference_leg(work.cpu_einsums, work.flops, sample_e, steps, warmup)
    config = make_config(work, E, max(1, n_gpus), "weak", False)
    config["reference_sample_elements_per_step"] = sample_e
    return {
        "impl": "reference", "metric": METRIC,
        "value": gf, "unit": "GFLOP/s", "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": work.dtype, "data": "synthetic",
        "config": config,
        "cpu_baseline": {
            "value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "sample": (f"{sample_e} elements per step"
                       + ("" if sample_e == E else f" (bounded sample of the labelled {E})")
                       + ": C/OpenMP restatement of the loop nest generate_loopy emits "
                         "(trivial schedule, -O3 -ffast-math -fopenmp); the reference's "
                         "loopy->pocl path cannot run in this image"),
        },
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def make_config(work: Work, E: int, world: int, scaling: str, flush: bool) -> dict:
    return {
        "workload": f"{work.descr}, {E} elements per GPU (BASELINE {work.baseline_cfg})",
        "elements_per_gpu": E,
        "elements_total": E * world,
        "flops_per_element": work.flops,
        "bytes_per_element": work.bytes,
        "l2_policy": work.l2_policy(E, flush),
        "parallelism": f"element axis sharded over {world} GPU(s) ({scaling} scaling), no collective",
    }


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


# ------------------------------------------------------------------ GPU arm --
class Bench:
    """Everything the GPU arm shares between workloads: queue, peaks, barrier, L2 flusher."""

    def __init__(self, rank: int, world: int, local_rank: int):
        import torch

        import feinsum_b200 as f
        from feinsum_b200 import _cabi
        from feinsum_b200.data import device_info

        self.torch, self.cabi = torch, _cabi
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dist = None
        torch.cuda.set_device(local_rank)
        if world > 1:
            import torch.distributed as dist_mod

            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        self.cq = f.CudaQueue(local_rank)
        self.dev = self.cq.torch_device
        # measured peaks of this board in this run: burst (micro-kernel of a few ms) and sustained
        # (after ~0.7 s of load the board sits at its power cap and the SM clock drops)
        self.peak = {}
        for dt, codes in (("f64", (3, 0)), ("f32", (1,))):
            self.peak[dt] = {
                "burst": max(_cabi.measure_peak(c) for c in codes),
                "sustained": max(_cabi.measure_peak(c + 16) for c in codes),
            }
            self.peak[dt]["sustained"] = min(self.peak[dt]["sustained"], self.peak[dt]["burst"])
        self.hbm_peak = device_info.DEV_TO_PEAK_BW.get("NVIDIA B200", 6561.6)
        self.hbm_src = "MEASURED_PEAKS.json" if device_info._hbm is not None else "fallback table"
        self.hbm_here = _cabi.measure_peak(2)      # this library's own streaming-copy micro-kernel
        self._flush_buf = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        torch = self.torch
        if self._flush_buf is None:
            self._flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        with torch.cuda.stream(self.cq.torch_stream):
            self._flush_buf.fill_(1)

    # ------------------------------------------------------------------
    def setup(self, work: Work, E: int):
        torch = self.torch
        gen = torch.Generator(device=self.dev).manual_seed(1234 + self.rank)
        tdt = torch.float32 if work.dtype == "f32" else torch.float64
        in_shapes, out_shapes = work.shapes(E)
        arrays = {n: torch.rand(s, dtype=tdt, device=self.dev, generator=gen) for n, s in in_shapes.items()}
        outs = {n: torch.zeros(s, dtype=tdt, device=self.dev) for n, s in out_shapes.items()}
        ex = work.program.executor(self.cq)
        return ex, arrays, outs

    def time_steps(self, ex, arrays, outs, steps: int, warmup: int, flush: bool, sample_clocks: bool):
        """(ms per step on this rank, launches per step, clocks summary)."""
        torch, cq = self.torch, self.cq
        for _ in range(warmup):
            ex(cq, **arrays, **outs)
        self.barrier()
        launches0 = self.cabi.launch_count()
        clocks = ClockSampler(self.local_rank) if sample_clocks else None
        if clocks is not None:
            clocks.__enter__()
        if flush:
            pairs = []
            for _ in range(steps):
                self.flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(cq.torch_stream)
                ex(cq, **arrays, **outs)
                b.record(cq.torch_stream)
                pairs.append((a, b))
            torch.cuda.synchronize()
            ms_total = sum(a.elapsed_time(b) for a, b in pairs)
        else:
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record(cq.torch_stream)
            for _ in range(steps):
                ex(cq, **arrays, **outs)
            stop.record(cq.torch_stream)
            stop.synchronize()
            ms_total = start.elapsed_time(stop)
        self.barrier()
        if clocks is not None:
            clocks.__exit__()
        launches = self.cabi.launch_count() - launches0
        return ms_total / steps, launches, (clocks.summary() if clocks is not None else None), ms_total

    def roofline(self, work: Work, E: int, my_ms: float, launches_per_step: int, long_region: bool) -> dict:
        pk = self.peak[work.dtype]
        ach_tflops = work.flops * E / (my_ms * 1e-3) * 1e-12
        ach_gbs = work.bytes * E / (my_ms * 1e-3) * 1e-9
        t_mem = work.bytes * E / (self.hbm_peak * 1e9)

        def frac_against(peak_fp):
            return max(work.flops * E / (peak_fp * 1e9), t_mem) / (my_ms * 1e-3)

        kind = "sustained" if long_region else "burst"
        peak_fp = pk[kind]
        t_flop = work.flops * E / (peak_fp * 1e9)
        if t_flop >= t_mem:
            roof = {"bound": "tensor", "achieved": ach_tflops, "peak": peak_fp * 1e-3,
                    "unit": "TFLOP/s", "frac": ach_tflops / (peak_fp * 1e-3)}
        else:
            roof = {"bound": "hbm", "achieved": ach_gbs, "peak": self.hbm_peak, "unit": "GB/s",
                    "frac": ach_gbs / self.hbm_peak}
        engine = "FP64 DMMA/DFMA" if work.dtype == "f64" else "FP32 FFMA2"
        roof.update({
            "peak_kind": kind + ("" if not long_region else " (timed region long enough to pull the power cap)"),
            "frac_burst": frac_against(pk["burst"]), "frac_sustained": frac_against(pk["sustained"]),
            "peak_burst": pk["burst"] * 1e-3, "peak_sustained": pk["sustained"] * 1e-3,
            "peak_source": (f"{engine} peaks measured in this run by fnsm_b200_measure_peak (burst: a few ms; "
                            f"sustained: after 0.7 s of load; MEASURED_PEAKS.json has no {work.dtype} figure); "
                            f"HBM {self.hbm_peak} GB/s from {self.hbm_src}"),
            "t_roof_ms": max(t_flop, t_mem) * 1e3, "roofline_frac": max(t_flop, t_mem) / (my_ms * 1e-3),
            "hbm": {"achieved": ach_gbs, "peak": self.hbm_peak, "frac": ach_gbs / self.hbm_peak,
                    "streaming_copy_measured_here": self.hbm_here,
                    "frac_of_best_measured": ach_gbs / max(self.hbm_peak, self.hbm_here)},
            "fp": {"achieved": ach_tflops, "peak": peak_fp * 1e-3, "frac": ach_tflops / (peak_fp * 1e-3)},
            "launches_per_step": launches_per_step,
            "algorithmic_bytes_per_launch": work.bytes * E / launches_per_step,
            "algorithmic_flops_per_launch": work.flops * E / launches_per_step,
        })
        return roof

    def run_workload(self, work: Work, E: int, steps: int, warmup: int, flush: bool = False,
                     sample_clocks: bool = False) -> dict:
        """Time one workload on every rank; returns the entry of the JSON line (same fields for the
        headline and for the suite)."""
        torch = self.torch
        ex, arrays, outs = self.setup(work, E)
        my_ms, launches, clocks, ms_total = self.time_steps(ex, arrays, outs, steps, warmup, flush, sample_clocks)
        capped = bool(clocks and "sw_power_cap" in clocks["reasons"])
        roof = self.roofline(work, E, my_ms, max(1, launches // steps), ms_total > 250.0 or capped)
        ms_step = self.max_over_ranks(my_ms)
        entry = {
            "elements_per_gpu": E, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
            "value": work.flops * E * self.world / (ms_step * 1e-3) * 1e-9, "unit": "GFLOP/s",
            "gbs": work.bytes * E * self.world / (ms_step * 1e-3) * 1e-9,
            "dtype": work.dtype, "roofline": roof, "gpu_launches": int(launches),
            "kernel": work.program.kernel_id, "l2_policy": work.l2_policy(E, flush),
            "baseline_config": work.baseline_cfg,
        }
        if clocks is not None:
            entry["clocks"] = clocks
        self._last = (ex, arrays, outs)
        del ex, arrays, outs
        return entry

    def release(self):
        self._last = None
        self.torch.cuda.empty_cache()

    # ---------------------------------------------------------------- e2e ---
    def e2e(self, work: Work, E: int, steps: int) -> dict:
        """The headline metric through the host-buffer API: pinned numpy in, pinned numpy out, the
        chunked H2D | kernel | D2H pipeline of HostExecutor inside the timed region; plus the raw
        pinned-copy time of the same bytes on the same ranks at the same time (the PCIe floor)."""
        from feinsum_b200.host_exec import HostExecutor, pinned_empty

        torch = self.torch
        ex, arrays, outs = self._last
        in_shapes, out_shapes = work.shapes(E)
        host_in = {}
        for name, shape in in_shapes.items():
            h = pinned_empty(shape, work.np_dtype)
            torch.from_numpy(h).copy_(arrays[name])
            host_in[name] = h
        host_out = {n: pinned_empty(s, work.np_dtype) for n, s in out_shapes.items()}
        hx = HostExecutor(work.program, self.cq)
        hx(outputs=host_out, **host_in)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            hx(outputs=host_out, **host_in)
        self.barrier()
        dt = self.max_over_ranks((time.perf_counter() - t0) / steps)
        # spot parity of the host path against the device path
        for oname in out_shapes:
            ref = outs[oname][..., :1].cpu().numpy()
            got = host_out[oname][..., :1]
            if not np.allclose(got, ref, rtol=1e-12 if work.dtype == "f64" else 1e-5):
                raise SystemExit("host path and device path disagree")
        # PCIe floor: the same bytes as plain contiguous pinned copies, H2D and D2H concurrently
        s_up, s_dn = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)
        dev_in = {n: torch.empty_like(arrays[n]) for n in arrays}

        def raw():
            with torch.cuda.stream(s_up):
                for n in dev_in:
                    dev_in[n].copy_(torch.from_numpy(host_in[n]), non_blocking=True)
            with torch.cuda.stream(s_dn):
                for n in host_out:
                    torch.from_numpy(host_out[n]).copy_(outs[n], non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()

        raw()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            raw()
        self.barrier()
        floor = self.max_over_ranks((time.perf_counter() - t0) / steps)
        return {"value": work.flops * E * self.world / dt * 1e-9, "unit": "GFLOP/s",
                "h2d_bytes_per_step": hx.h2d_bytes, "d2h_bytes_per_step": hx.d2h_bytes,
                "ms_per_step": dt * 1e3, "steps": steps,
                "pcie_floor_ms": floor * 1e3, "pcie_frac": floor / dt,
                "h2d_gbs_per_gpu": hx.h2d_bytes / dt * 1e-9, "pcie_floor_h2d_gbs_per_gpu": hx.h2d_bytes / floor * 1e-9,
                "note": "HostExecutor: pinned numpy in/out, chunked H2D | kernel | D2H on 3 streams; "
                        "pcie_floor_ms = the same bytes as plain pinned cudaMemcpyAsync (H2D and D2H "
                        "concurrently, all ranks at once), pcie_frac = floor / e2e time"}


def cpu_baseline_subprocess(workload: str, E: int, steps: int = 5, warmup: int = 2) -> dict | None:
    """Runs this file's reference arm in a fresh process (no torch / CUDA state, all host cores) so the
    GPU arm's cpu_baseline and `--impl reference` are the same code under the same policy."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("OMP_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        out = subprocess.run(
            [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
             "--elements", str(E), "--steps", str(steps), "--warmup", str(warmup)],
            capture_output=True, text=True, env=env, timeout=600, check=True).stdout
        line = json.loads(out.strip().splitlines()[-1])
        cpu = line["cpu_baseline"]
        cpu["sample"] += f"; {line['steps']} steps after {line['warmup']} warm-ups, {line['ms_per_step']:.1f} ms/step"
        return cpu
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "unit": "GFLOP/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                "sample": f"failed: {exc}"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="div_p4", choices=sorted(WORKLOADS))
    ap.add_argument("--elements", type=int, default=0, help="elements per GPU (default: workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: the headline runs a fixed total of 16 M elements split over the ranks")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 dmma, 2 simt")
    ap.add_argument("--param", action="append", default=[], metavar="K=V",
                    help="launch parameter of the kernel (threads=384, ...); may repeat")
    ap.add_argument("--flush-l2", action="store_true", help="flush L2 between timed launches of the headline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-suite", action="store_true", help="headline workload only")
    ap.add_argument("--suite-steps", type=int, default=10)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    params = {k: int(v) for k, v in (kv.split("=") for kv in args.param)}
    if args.variant > 0:
        params["variant"] = args.variant
    work = Work(args.workload, params)
    if args.scaling == "strong":
        E = (args.elements or STRONG_TOTAL) // max(world, 1)
    else:
        E = args.elements or work.default_e

    # ------------------------------------------------------ reference arm ---
    if args.impl == "reference":
        if rank != 0:
            return
        print(json.dumps(reference_arm(work, E, args.gpus, args.steps, args.warmup, None)))
        return

    # ------------------------------------------------------------ B200 arm --
    bench = Bench(rank, world, local_rank)
    custom = bool(params) or bool(args.elements) or args.workload != "div_p4" or args.scaling != "weak"
    head = bench.run_workload(work, E, args.steps, args.warmup, flush=args.flush_l2, sample_clocks=True)
    traffic, traffic_src = (profiled_traffic(args.workload)
                            if E == work.default_e and not params else (None, None))
    head["roofline"].update({"traffic": traffic, "traffic_source": traffic_src})
    e2e = None
    if not args.no_e2e:
        e2e = bench.e2e(work, E, max(3, min(args.steps, 5)))
    bench.release()

    suite, strong = {}, {}
    if not args.no_suite and not custom:
        for name, n, flush in SUITE:
            w = Work(name)
            n = n or w.default_e
            key = name + (f"@{n}" if n != w.default_e else "") + ("+l2flush" if flush else "")
            # launches of tens of microseconds: time enough of them that the start of the loop (an empty queue, no
            # previous kernel to overlap the launch with) and the clock sampler's subprocess do not show
            steps = args.suite_steps * (10 if n <= 500_000 and not flush else 1)
            suite[key] = bench.run_workload(w, n, steps, 3 if steps == args.suite_steps else 10, flush=flush,
                                            sample_clocks=True)
            bench.release()
        for name in STRONG_SUITE:
            w = Work(name)
            n = STRONG_TOTAL // world
            ent = bench.run_workload(w, n, args.suite_steps, 3, sample_clocks=True)
            ent["elements_total"] = n * world
            strong[name] = ent
            bench.release()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_subprocess(args.workload, E)

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": head["value"], "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": work.dtype, "data": "synthetic",
            "config": make_config(work, E, world, args.scaling, args.flush_l2),
            "gbs": head["gbs"], "roofline": head["roofline"], "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": head["gpu_launches"] + sum(v["gpu_launches"] for v in suite.values())
                            + sum(v["gpu_launches"] for v in strong.values()),
            "gpu_launches_headline": head["gpu_launches"],
            "clocks": head.get("clocks"), "kernel": head["kernel"], "device": bench.cq.device.name,
        }
        if suite:
            line["workloads"] = suite
        if strong:
            line["strong"] = {"elements_total": STRONG_TOTAL, "note": "fixed total split over the ranks; "
                              "speed-up at N = value_N / value_1 of the same entry", **strong}
        print(json.dumps(line))
    if bench.dist is not None:
        bench.dist.destroy_process_group()


if __name__ == "__main__":
    main()
