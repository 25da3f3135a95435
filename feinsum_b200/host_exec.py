"""
Host-buffer execution: the call a user with numpy arrays makes.

``HostExecutor(program, cq)(**numpy_arrays) -> {"_fe_out": ndarray, ...}``
moves the operands to the B200, runs the kernel and brings the results back,
pipelined over chunks of the element axis so that host->device copies, the
kernel and device->host copies of neighbouring chunks overlap on three CUDA
streams (PCIe is full duplex; the kernel is ~50x faster than either copy, so
the end-to-end time is max(H2D, D2H) plus one chunk of latency).

Operands whose element axis is not the leading one (``J(3,3,E)``,
``u(3,E,35)`` ...) are not contiguous per chunk; they are moved with strided
2-D copies (``fnsm_b200_copy2d_async`` = ``cudaMemcpy2DAsync``), never by
re-packing on the host.  Host arrays should be page-locked
(:func:`pinned_empty`) -- pageable memory works but serialises the copies.

*program* is a :class:`~feinsum_b200.codegen.cuda.CudaProgram` or anything with
the same three members (``host_spec()``, ``executor(cq)``), e.g.
:class:`feinsum_b200.wave3d.Wave3DProgram` -- the three-einsum wave operator
goes through the same chunk pipeline.

This is the reference-facing path that ``bench.py`` reports as ``e2e``.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any

import numpy as np

from feinsum_b200 import _cabi
from feinsum_b200.cl_utils import CudaQueue, as_queue
from feinsum_b200.einsum import BatchedEinsum, SizeParam


def pinned_empty(shape: tuple[int, ...], dtype: Any) -> np.ndarray:
    """Page-locked host array: a numpy view of a pinned torch tensor.  The view's ``base`` chain
    owns the tensor, so the pinned allocation is released when the last view of it goes away."""
    import torch

    tdt = {np.dtype("float64"): torch.float64, np.dtype("float32"): torch.float32}[np.dtype(dtype)]
    return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()


@dataclass(frozen=True)
class HostSpec:
    """Names, symbolic shapes and dtypes of what crosses the host boundary."""

    in_shapes: dict[str, tuple[Any, ...]]
    in_dtypes: dict[str, np.dtype[Any]]
    out_shapes: dict[str, tuple[Any, ...]]
    out_dtypes: dict[str, np.dtype[Any]]


def einsum_host_spec(es: BatchedEinsum) -> HostSpec:
    return HostSpec(
        {n: tuple(s) for n, s in es.arg_to_shape.items()},
        {n: np.dtype(d) for n, d in es.arg_to_dtype.items()},
        {n: tuple(es.shape) for n in es.output_names},
        {n: np.dtype(np.result_type(*[a.dtype for a in row]))
         for n, row in zip(es.output_names, es.args)},
    )


def _long_axis(shape: tuple[Any, ...]) -> int | None:
    axes = [k for k, d in enumerate(shape) if isinstance(d, SizeParam)]
    if len(axes) > 1:
        raise NotImplementedError("operands with two symbolic axes are not chunked")
    return axes[0] if axes else None


class HostExecutor:
    def __init__(self, program: Any, cq: Any = None, chunk: int = 262144):
        import torch

        self.program = program
        self.spec: HostSpec = (program.host_spec() if hasattr(program, "host_spec")
                               else einsum_host_spec(program.einsum))
        self.cq: CudaQueue = as_queue(cq)
        self.chunk = int(chunk)
        if self.chunk < 1:
            raise ValueError("chunk must be positive")
        self.lib = _cabi.lib()
        dev = self.cq.torch_device
        self._streams = {k: torch.cuda.Stream(device=dev) for k in ("h2d", "run", "d2h")}
        self._run_q = CudaQueue(self.cq.device, self._streams["run"])
        self._exec = program.executor(self._run_q)
        self._bufs: dict[Any, Any] = {}
        self._const: dict[str, Any] = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ------------------------------------------------------------------
    def _dev_buf(self, key: Any, shape: tuple[int, ...], dtype: np.dtype[Any]) -> Any:
        import torch

        tdt = torch.float64 if np.dtype(dtype) == np.dtype("float64") else torch.float32
        buf = self._bufs.get(key)
        if buf is None or tuple(buf.shape) != shape or buf.dtype != tdt:
            buf = torch.empty(shape, dtype=tdt, device=self.cq.torch_device)
            self._bufs[key] = buf
        return buf

    def _copy(self, dst_ptr: int, dpitch: int, src_ptr: int, spitch: int, width: int,
              height: int, kind: int, stream: Any) -> None:
        rc = self.lib.fnsm_b200_copy2d_async(
            C.c_void_p(dst_ptr), dpitch, C.c_void_p(src_ptr), spitch, width, height, kind,
            C.c_void_p(int(stream.cuda_stream)),
        )
        _cabi.check(rc, "fnsm_b200_copy2d_async")

    # ------------------------------------------------------------------
    def __call__(self, outputs: dict[str, np.ndarray] | None = None, **arrays: np.ndarray) -> dict[str, np.ndarray]:
        import torch

        spec = self.spec
        sizes: dict[str, int] = {}
        for name, shape in spec.in_shapes.items():
            if name not in arrays:
                raise TypeError(f"missing input array '{name}'")
            a = arrays[name]
            if not isinstance(a, np.ndarray) or not a.flags.c_contiguous:
                raise TypeError(f"'{name}' must be a C-contiguous numpy array")
            if a.dtype != spec.in_dtypes[name] or a.ndim != len(shape):
                raise TypeError(f"'{name}' has wrong dtype or rank")
            for d, got in zip(shape, a.shape):
                if isinstance(d, SizeParam):
                    if sizes.setdefault(d.name, int(got)) != int(got):
                        raise ValueError(f"inconsistent size parameter '{d.name}'")
                elif int(d) != int(got):
                    raise ValueError(f"'{name}': expected {shape}, got {a.shape}")
        unknown = set(arrays) - set(spec.in_shapes)
        if unknown:
            raise TypeError(f"unexpected arguments: {sorted(unknown)}")
        if len(sizes) > 1:
            raise NotImplementedError("one symbolic axis expected")
        E = next(iter(sizes.values())) if sizes else 1
        out_axes = {n: _long_axis(s) for n, s in spec.out_shapes.items()}
        out_shapes = {n: tuple(E if isinstance(d, SizeParam) else int(d) for d in s)
                      for n, s in spec.out_shapes.items()}
        outs = dict(outputs) if outputs else {}
        for oname, odt in spec.out_dtypes.items():
            if oname not in outs:
                outs[oname] = pinned_empty(out_shapes[oname], odt)
            elif outs[oname].shape != out_shapes[oname] or outs[oname].dtype != odt \
                    or not outs[oname].flags.c_contiguous:
                raise ValueError(f"output '{oname}' has wrong shape, dtype or layout")
        chunked = bool(sizes) and all(ax is not None for ax in out_axes.values())

        s_h2d, s_run, s_d2h = (self._streams[k] for k in ("h2d", "run", "d2h"))
        self.h2d_bytes = self.d2h_bytes = 0
        with torch.cuda.device(self.cq.torch_device):
            # operands without the element axis: one copy, before everything else
            for name, shape in spec.in_shapes.items():
                if _long_axis(shape) is None or not chunked:
                    a = arrays[name]
                    buf = self._dev_buf(("const", name), tuple(a.shape), a.dtype)
                    if a.nbytes:
                        self._copy(buf.data_ptr(), a.nbytes, a.ctypes.data, a.nbytes, a.nbytes, 1, 0, s_h2d)
                    self.h2d_bytes += a.nbytes
                    self._const[name] = buf
            if not chunked:
                chunks = [(0, E)]          # nothing to chunk over: single shot
            else:
                chunks = [(s, min(E, s + self.chunk)) for s in range(0, E, self.chunk)]
            ev_h2d = [None, None]
            ev_run = [None, None]
            ev_d2h = [None, None]
            for ci, (lo, hi) in enumerate(chunks):
                slot = ci & 1
                n = hi - lo
                # ---- H2D (inputs of chunk ci into buffer set `slot`)
                if ev_run[slot] is not None:
                    s_h2d.wait_event(ev_run[slot])
                dev_in: dict[str, Any] = {}
                for name, shape in spec.in_shapes.items():
                    ax = _long_axis(shape)
                    a = arrays[name]
                    if ax is None or not chunked:
                        dev_in[name] = self._const[name]
                        continue
                    cshape = tuple(n if k == ax else int(a.shape[k]) for k in range(a.ndim))
                    buf = self._dev_buf((slot, name, n), cshape, a.dtype)
                    inner = int(np.prod(a.shape[ax + 1:], dtype=np.int64)) * a.itemsize
                    outer = int(np.prod(a.shape[:ax], dtype=np.int64))
                    self._copy(buf.data_ptr(), n * inner, a.ctypes.data + lo * inner,
                               a.shape[ax] * inner, n * inner, outer, 0, s_h2d)
                    self.h2d_bytes += outer * n * inner
                    dev_in[name] = buf
                ev_h2d[slot] = torch.cuda.Event()
                ev_h2d[slot].record(s_h2d)
                # ---- kernel
                s_run.wait_event(ev_h2d[slot])
                if ev_d2h[slot] is not None:
                    s_run.wait_event(ev_d2h[slot])
                dev_out = {}
                for oname, odt in spec.out_dtypes.items():
                    oshape, oax = out_shapes[oname], out_axes[oname]
                    cshape = tuple(n if (chunked and k == oax) else oshape[k] for k in range(len(oshape)))
                    dev_out[oname] = self._dev_buf((slot, oname, n), cshape, odt)
                self._exec(self._run_q, **dev_in, **dev_out)
                ev_run[slot] = torch.cuda.Event()
                ev_run[slot].record(s_run)
                # ---- D2H
                s_d2h.wait_event(ev_run[slot])
                for oname in spec.out_dtypes:
                    o, buf, oax = outs[oname], dev_out[oname], out_axes[oname]
                    if not o.nbytes:
                        continue
                    if not chunked:
                        self._copy(o.ctypes.data, o.nbytes, buf.data_ptr(), o.nbytes, o.nbytes, 1, 1, s_d2h)
                        self.d2h_bytes += o.nbytes
                    else:
                        inner = int(np.prod(o.shape[oax + 1:], dtype=np.int64)) * o.itemsize
                        outer = int(np.prod(o.shape[:oax], dtype=np.int64))
                        self._copy(o.ctypes.data + lo * inner, o.shape[oax] * inner,
                                   buf.data_ptr(), n * inner, n * inner, outer, 1, s_d2h)
                        self.d2h_bytes += outer * n * inner
                ev_d2h[slot] = torch.cuda.Event()
                ev_d2h[slot].record(s_d2h)
            for s in (s_h2d, s_run, s_d2h):
                s.synchronize()
        return outs
