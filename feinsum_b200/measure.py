"""
Validation, timing and roofline reporting of batched einsums on a B200.

Same entry points and semantics as the reference's ``feinsum.measure``
(reference ``src/feinsum/measure.py:35-525``), rebuilt on CUDA:

* :func:`validate_batched_einsum_transform` -- run the kernel at E = 100 on
  ``default_rng(0)`` inputs and compare every output with
  ``numpy.einsum(..., optimize="optimal")`` to 1e-10 (fp64) / 1e-6 (fp32)
  (reference ``measure.py:111-194``).
* :func:`timeit` -- validate first, then 5 warm-up launches and batches of 5
  launches until >= 10 launches and >= 2 s have been timed (reference
  ``measure.py:35-37,250-275``); the clock is a pair of CUDA events on the
  launching stream instead of host ``time()``.
* :func:`measure_giga_op_rate`, :func:`get_roofline_flop_rate`,
  :func:`stringify_comparison_vs_roofline` -- the reference's op-count and
  footprint models (``measure.py:278-418``), against B200 peaks.

FLOP model: for every step of the flop-optimal contraction schedule,
``prod(extents of all indices of the step) * ((#operands - 1) + [step sums])``
per output row -- loopy's op-map count of the scheduled kernel (one per
multiply, one per reduction add; pinned by the reference's tests at 33 075 /
7 980 flops per element for the trivial / hoisted DG gradient,
``test/test_loopy_utils.py:267-271``).  Byte model: every distinct input and
every output counted once (``measure.py:334-354``).
"""

from __future__ import annotations

import logging
from collections.abc import Mapping
from typing import Any

import numpy as np

from feinsum_b200._immutable import Map
from feinsum_b200.cl_utils import as_queue
from feinsum_b200.codegen.cuda import generate_cuda
from feinsum_b200.contraction_schedule import (
    ContractionSchedule,
    get_opt_einsum_contraction_schedule,
)
from feinsum_b200.diagnostics import NoDevicePeaksInfoError, TransformValidationError
from feinsum_b200.einsum import INT_CLASSES, BatchedEinsum, IntegralT, SizeParam
from feinsum_b200.make_einsum import parse_subscripts

logger = logging.getLogger(__name__)

FP32_VALIDATION_TOL = 5e-6
N_WARMUP_ROUNDS = 5
N_MIN_TIMING_ROUNDS = 10
N_MIN_SIM_SECS = 2


def fp32_validation_tol(program: Any) -> float:
    """fp32 acceptance tolerance of :func:`validate_batched_einsum_transform`.

    The reference's gate is ``atol = rtol = 1e-6`` against numpy's own fp32 evaluation
    (reference ``measure.py:178-180``); it is kept for every kernel that accumulates in plain
    fp32 (simt and generic variants, tensor-product).  The DG operator kernels' default fp32 path
    is 3xTF32 on the tensor cores (``variant`` 0/1/3): split operands, fp32 accumulation with
    truncation -- measured worst case 1.2e-6 against the fp64 oracle (div), plus numpy's own
    ~4e-7.  For those variants only, the gate is :data:`FP32_VALIDATION_TOL` = 5e-6, half of the
    1e-5 the drop-in contract allows for fp32.  (Behavioural difference vs the reference;
    INTEGRATION.md section "fp32 validation gate".)"""
    plan = getattr(program, "plan", None)
    if plan is not None and plan.kernel_id in ("grad", "div", "lift_ef", "lift_fe", "opmat_se") \
            and int(dict(getattr(program, "params", {})).get("variant", 0)) != 2:
        return FP32_VALIDATION_TOL
    return 1e-6


def get_real_dtype(dtype: np.dtype[Any]) -> np.dtype[Any]:
    return np.empty(0, dtype=dtype).real.dtype


# {{{ inputs / outputs


def _generate_random_np_array(
    rng: np.random.Generator, dtype: np.dtype[Any], shape: tuple[IntegralT, ...]
) -> np.ndarray:
    # reference measure.py:63-77
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        real = get_real_dtype(dtype)
        return rng.random(size=shape, dtype=real) + dtype.type(1j) * rng.random(
            size=shape, dtype=real
        )
    if dtype.kind == "i":
        return rng.integers(low=-100, high=100, size=shape, dtype=dtype)
    return rng.random(size=shape, dtype=dtype)


def _concrete_shape(shape: tuple[Any, ...], long_dim_length: int) -> tuple[int, ...]:
    return tuple(
        int(d) if isinstance(d, INT_CLASSES) else int(long_dim_length) for d in shape
    )


def generate_host_input_arrays(
    einsum: BatchedEinsum, long_dim_length: int, np_seed: int = 0
) -> dict[str, np.ndarray]:
    """numpy inputs, drawn in sorted-operand-name order (deterministic; the
    reference's order depends on string hashing, ``measure.py:101-108``)."""
    rng = np.random.default_rng(np_seed)
    return {
        name: _generate_random_np_array(
            rng,
            einsum.arg_to_dtype[name],
            _concrete_shape(einsum.arg_to_shape[name], long_dim_length),
        )
        for name in sorted(einsum.arg_to_dtype)
    }


def generate_input_arrays(
    queue: Any, einsum: BatchedEinsum, long_dim_length: int, np_seed: int = 0
) -> Map[str, Any]:
    """Device inputs for every operand (reference ``measure.py:80-108``)."""
    import torch

    q = as_queue(queue)
    host = generate_host_input_arrays(einsum, long_dim_length, np_seed)
    return Map({k: torch.from_numpy(v).to(q.torch_device) for k, v in host.items()})


def generate_out_arrays(
    queue: Any, einsum: BatchedEinsum, long_dim_length: int
) -> Map[str, Any]:
    """Zero-initialised outputs (reference ``measure.py:43-60``)."""
    import torch

    q = as_queue(queue)
    shape = _concrete_shape(einsum.shape, long_dim_length)
    outs = {}
    for name, row in zip(einsum.output_names, einsum.args):
        dt = np.result_type(*[a.dtype for a in row])
        outs[name] = torch.zeros(shape, dtype=getattr(torch, np.dtype(dt).name), device=q.torch_device)
    return Map(outs)


# }}}


def _lower(einsum: BatchedEinsum, transform: Any, schedule: ContractionSchedule | None) -> Any:
    program = generate_cuda(einsum, schedule=schedule)
    if transform is not None:
        program = transform(program, insn_match=None, kernel_name=None)
    return program


def validate_batched_einsum_transform(
    einsum: BatchedEinsum,
    cq: Any,
    transform: Any,
    schedule: ContractionSchedule | None = None,
) -> None:
    """Raise :class:`TransformValidationError` if the configured kernel does
    not reproduce ``numpy.einsum`` (reference ``measure.py:111-194``)."""
    q = as_queue(cq)
    long_dim_length = 100
    program = _lower(einsum, transform, schedule)
    arg_dict = dict(generate_input_arrays(q, einsum, long_dim_length))

    subscripts = einsum.get_subscripts()
    ref_outs = {
        name: np.einsum(
            subscripts,
            *[arg_dict[arg.name].cpu().numpy() for arg in row],
            optimize="optimal",
        )
        for name, row in zip(einsum.output_names, einsum.args)
    }

    executor = program.executor(q, **arg_dict)
    evt, outs = executor(q, **arg_dict)
    evt.wait()

    if frozenset(ref_outs) != frozenset(outs):
        raise RuntimeError("Output names mismatch")

    for name, ref_out in sorted(ref_outs.items()):
        got = outs[name].cpu().numpy()
        if ref_out.dtype != got.dtype:
            raise RuntimeError(f"dtype mismatch for output '{name}'")
        real = get_real_dtype(ref_out.dtype)
        if real == np.float32:
            atol = rtol = fp32_validation_tol(program)
        elif real == np.float64:
            atol = rtol = 1e-10
        else:
            raise NotImplementedError(real)
        try:
            np.testing.assert_allclose(got, ref_out, atol=atol, rtol=rtol)
        except AssertionError as exc:
            raise TransformValidationError(f"{exc}") from exc

    logger.info("Statistically verified the soundness of the transformation")


def time_executor(
    executor: Any,
    q: Any,
    arg_dict: Mapping[str, Any],
    *,
    warmup: int = N_WARMUP_ROUNDS,
    min_rounds: int = N_MIN_TIMING_ROUNDS,
    min_secs: float = N_MIN_SIM_SECS,
    batch: int = 5,
) -> float:
    """Average seconds per launch, CUDA events on the launching stream."""
    import torch

    for _ in range(warmup):
        executor(q, **arg_dict)
    q.finish()

    total_time, total_rounds = 0.0, 0
    while total_rounds < min_rounds or total_time < min_secs:
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record(q.torch_stream)
        for _ in range(batch):
            executor(q, **arg_dict)
        stop.record(q.torch_stream)
        stop.synchronize()
        total_time += start.elapsed_time(stop) * 1e-3
        total_rounds += batch
    return total_time / total_rounds


def timeit(
    einsum: BatchedEinsum,
    *,
    transform: Any,
    cq: Any,
    long_dim_length: int = 100000,
    schedule: ContractionSchedule | None = None,
) -> float:
    """Seconds per execution of *einsum* on *cq*'s device (reference
    ``measure.py:197-275``; device-event clock instead of host wall clock)."""
    q = as_queue(cq)
    validate_batched_einsum_transform(einsum, q, transform, schedule)

    program = _lower(einsum, transform, schedule)
    arg_dict = dict(generate_input_arrays(q, einsum, long_dim_length))
    arg_dict.update(generate_out_arrays(q, einsum, long_dim_length))
    executor = program.executor(q, **arg_dict)
    return time_executor(
        executor, q, arg_dict,
        warmup=N_WARMUP_ROUNDS, min_rounds=N_MIN_TIMING_ROUNDS, min_secs=N_MIN_SIM_SECS,
    )


# {{{ op / byte model


def _step_flops(subscripts: str, extent: Mapping[str, float]) -> float:
    out_idx, in_idx_sets = parse_subscripts(subscripts)
    touched: set[str] = set()
    for s in in_idx_sets:
        touched.update(s)
    iters = 1.0
    for idx in touched:
        iters *= extent[idx]
    ops = len(in_idx_sets) - 1 + (1 if touched - set(out_idx) else 0)
    return iters * ops


def get_flops_per_dtype(
    einsum: BatchedEinsum,
    long_dim_length: int,
    schedule: ContractionSchedule | None = None,
) -> Map[np.dtype[Any], float]:
    """FLOPs of the (flop-optimal unless given) schedule, by result dtype."""
    if schedule is None:
        schedule = get_opt_einsum_contraction_schedule(einsum)
    flops: dict[np.dtype[Any], float] = {}
    # extents of step-local indices == the einsum's (intermediates reuse letters)
    extent = {
        idx: float(long_dim_length) if isinstance(ext, SizeParam) else float(ext)
        for idx, ext in einsum.index_to_dim_length.items()
    }
    per_row = sum(_step_flops(s, extent) for s in schedule.subscripts)
    for row in einsum.args:
        dt = np.dtype(np.result_type(*[a.dtype for a in row]))
        weight = 1.0
        if dt.kind == "c":
            # reference weights complex add/mul 2/6 (measure.py:307-320); a
            # multiply-add pair therefore counts 8 real flops per 2 ops
            weight, dt = 4.0, get_real_dtype(dt)
        flops[dt] = flops.get(dt, 0.0) + weight * per_row
    return Map(flops)


def _get_giga_ops_from_einsum(
    expr: BatchedEinsum, long_dim_length: int = 100_000
) -> Map[np.dtype[Any], float]:
    return Map(
        {k: v * 1e-9 for k, v in get_flops_per_dtype(expr, long_dim_length).items()}
    )


def get_footprint_bytes(expr: BatchedEinsum, long_dim_length: int) -> float:
    """Compulsory traffic: each distinct input and each output once."""
    total = 0.0
    for name, shape in expr.arg_to_shape.items():
        total += float(np.prod(_concrete_shape(shape, long_dim_length), dtype=np.float64)) * np.dtype(
            expr.arg_to_dtype[name]
        ).itemsize
    out_elems = float(np.prod(_concrete_shape(expr.shape, long_dim_length), dtype=np.float64))
    for row in expr.args:
        total += out_elems * np.dtype(np.result_type(*[a.dtype for a in row])).itemsize
    return total


def _get_footprint_gbytes(expr: BatchedEinsum, long_dim_length: int) -> float:
    return get_footprint_bytes(expr, long_dim_length) * 1e-9


# }}}


def measure_giga_op_rate(
    expr: BatchedEinsum,
    *,
    transform: Any,
    cq: Any,
    long_dim_length: int = 100000,
    schedule: ContractionSchedule | None = None,
) -> Map[np.dtype[Any], float]:
    """GOp/s by result dtype (reference ``measure.py:357-385``)."""
    runtime = timeit(
        expr, transform=transform, cq=cq, long_dim_length=long_dim_length, schedule=schedule
    )
    return Map(
        {k: v / runtime for k, v in _get_giga_ops_from_einsum(expr, long_dim_length).items()}
    )


def get_roofline_flop_rate(
    expr: BatchedEinsum, dev_name: str, long_dim_length: int = 100_000
) -> Map[np.dtype[Any], float]:
    """``ops / max(ops/peak_flops, bytes/peak_bw)`` (reference ``measure.py:388-418``)."""
    from feinsum_b200.data.device_info import DEV_TO_PEAK_BW, DEV_TO_PEAK_GFLOPS

    dtype_to_gflops = _get_giga_ops_from_einsum(expr, long_dim_length)
    ngbs = _get_footprint_gbytes(expr, long_dim_length)
    try:
        t_flops = max(
            ngflops / DEV_TO_PEAK_GFLOPS[dev_name][dtype.name]
            for dtype, ngflops in dtype_to_gflops.items()
        )
        t_bw = ngbs / DEV_TO_PEAK_BW[dev_name]
    except KeyError as exc:
        raise NoDevicePeaksInfoError from exc
    roofline_time = max(t_flops, t_bw)
    return Map({dtype: g / roofline_time for dtype, g in dtype_to_gflops.items()})


def _strify_measured_vs_roofline(
    measured: Mapping[np.dtype[Any], Any], roofline: Mapping[np.dtype[Any], Any]
) -> str:
    from tabulate import tabulate

    assert set(measured.keys()) == set(roofline.keys())
    table = [["Dtype", "Measured GOps/s", "Roofline GOps/s"]]
    for dtype in sorted(measured.keys(), key=lambda x: x.itemsize):
        m, r = measured[dtype], roofline[dtype]
        table.append(
            [
                dtype.name,
                f"{m:.1f}" if isinstance(m, float) else str(m),
                f"{r:.1f}" if isinstance(r, float) else str(r),
            ]
        )
    return tabulate(table, tablefmt="fancy_grid")


def _stringify_runtime_comparison_vs_roofline(
    expr: BatchedEinsum,
    runtime: float,
    device_name: str,
    *,
    long_dim_length: int = 100000,
    ignore_unknown_device: bool = True,
) -> str:
    measured = Map(
        {k: v / runtime for k, v in _get_giga_ops_from_einsum(expr, long_dim_length).items()}
    )
    try:
        roofline = get_roofline_flop_rate(expr, device_name, long_dim_length)
    except NoDevicePeaksInfoError:
        if ignore_unknown_device:
            return _strify_measured_vs_roofline(measured, dict.fromkeys(measured, "N/A"))
        raise
    return _strify_measured_vs_roofline(measured, roofline)


def stringify_comparison_vs_roofline(
    expr: BatchedEinsum,
    *,
    schedule: ContractionSchedule | None = None,
    transform: Any,
    cq: Any,
    long_dim_length: int = 100000,
    ignore_unknown_device: bool = True,
) -> str:
    """Pretty table ``Dtype | Measured GOps/s | Roofline GOps/s`` (reference
    ``measure.py:484-525``).  Unlike the reference (which always evaluates the
    roofline at E = 100 000) the roofline uses the
    timed ``long_dim_length``."""
    q = as_queue(cq)
    measured = measure_giga_op_rate(
        expr, transform=transform, schedule=schedule, cq=q, long_dim_length=long_dim_length
    )
    try:
        roofline = get_roofline_flop_rate(expr, q.device.name, long_dim_length)
    except NoDevicePeaksInfoError:
        if ignore_unknown_device:
            return _strify_measured_vs_roofline(measured, dict.fromkeys(measured, "N/A"))
        raise
    return _strify_measured_vs_roofline(measured, roofline)


def roofline_report(
    expr: BatchedEinsum, runtime: float, dev_name: str, long_dim_length: int
) -> dict[str, Any]:
    """GFLOP/s, GB/s and roofline fraction of one timed execution (BASELINE metric)."""
    from feinsum_b200.data.device_info import DEV_TO_PEAK_BW, DEV_TO_PEAK_GFLOPS

    flops = get_flops_per_dtype(expr, long_dim_length)
    nbytes = get_footprint_bytes(expr, long_dim_length)
    total_flops = sum(flops.values())
    rep: dict[str, Any] = {
        "gflops": total_flops / runtime * 1e-9,
        "gbs": nbytes / runtime * 1e-9,
        "flops": total_flops,
        "bytes": nbytes,
        "seconds": runtime,
    }
    if dev_name in DEV_TO_PEAK_BW and dev_name in DEV_TO_PEAK_GFLOPS:
        t_mem = nbytes * 1e-9 / DEV_TO_PEAK_BW[dev_name]
        t_flop = max(
            f * 1e-9 / DEV_TO_PEAK_GFLOPS[dev_name][dt.name] for dt, f in flops.items()
        )
        rep["t_roof"] = max(t_mem, t_flop)
        rep["bound"] = "hbm" if t_mem >= t_flop else "fp"
        rep["roofline_frac"] = rep["t_roof"] / runtime
    return rep
