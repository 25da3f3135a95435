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

This is the reference-facing path that ``bench.py`` reports as ``e2e``.
"""

from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np

from feinsum_b200 import _cabi
from feinsum_b200.cl_utils import CudaQueue, as_queue
from feinsum_b200.codegen.cuda import CudaProgram
from feinsum_b200.einsum import SizeParam


def pinned_empty(shape: tuple[int, ...], dtype: Any) -> np.ndarray:
    """Page-locked host array (numpy view of a pinned torch tensor)."""
    import torch

    tdt = {np.dtype("float64"): torch.float64, np.dtype("float32"): torch.float32}[np.dtype(dtype)]
    t = torch.empty(shape, dtype=tdt, pin_memory=True)
    a = t.numpy()
    _KEEPALIVE[id(a)] = t
    return a


_KEEPALIVE: dict[int, Any] = {}


def _long_axis(shape: tuple[Any, ...]) -> int | None:
    axes = [k for k, d in enumerate(shape) if isinstance(d, SizeParam)]
    if len(axes) > 1:
        raise NotImplementedError("operands with two symbolic axes are not chunked")
    return axes[0] if axes else None


class HostExecutor:
    def __init__(self, program: CudaProgram, cq: Any = None, chunk: int = 262144):
        import torch

        self.program = program
        self.einsum = program.einsum
        self.cq: CudaQueue = as_queue(cq)
        self.chunk = int(chunk)
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

        es = self.einsum
        sizes: dict[str, int] = {}
        for name, shape in es.arg_to_shape.items():
            a = arrays[name]
            if not isinstance(a, np.ndarray) or not a.flags.c_contiguous:
                raise TypeError(f"'{name}' must be a C-contiguous numpy array")
            if a.dtype != np.dtype(es.arg_to_dtype[name]) or a.ndim != len(shape):
                raise TypeError(f"'{name}' has wrong dtype or rank")
            for d, got in zip(shape, a.shape):
                if isinstance(d, SizeParam):
                    if sizes.setdefault(d.name, int(got)) != int(got):
                        raise ValueError(f"inconsistent size parameter '{d.name}'")
                elif int(d) != int(got):
                    raise ValueError(f"'{name}': expected {shape}, got {a.shape}")
        if len(sizes) > 1:
            raise NotImplementedError("one symbolic axis expected")
        out_axis = _long_axis(es.shape)
        E = next(iter(sizes.values())) if sizes else 1
        out_shape = tuple(E if isinstance(d, SizeParam) else int(d) for d in es.shape)
        out_dtypes = [np.dtype(np.result_type(*[a.dtype for a in row])) for row in es.args]
        outs = outputs or {}
        for oname, odt in zip(es.output_names, out_dtypes):
            if oname not in outs:
                outs[oname] = pinned_empty(out_shape, odt)
            elif outs[oname].shape != out_shape or outs[oname].dtype != odt:
                raise ValueError(f"output '{oname}' has wrong shape or dtype")

        s_h2d, s_run, s_d2h = (self._streams[k] for k in ("h2d", "run", "d2h"))
        self.h2d_bytes = self.d2h_bytes = 0
        with torch.cuda.device(self.cq.torch_device):
            # operands without the element axis: one copy, before everything else
            for name, shape in es.arg_to_shape.items():
                if _long_axis(shape) is None:
                    a = arrays[name]
                    buf = self._dev_buf(("const", name), tuple(a.shape), a.dtype)
                    self._copy(buf.data_ptr(), a.nbytes, a.ctypes.data, a.nbytes, a.nbytes, 1, 0, s_h2d)
                    self.h2d_bytes += a.nbytes
                    self._const[name] = buf
            if not sizes or out_axis is None:
                # nothing to chunk over: single shot
                chunks = [(0, E)]
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
                for name, shape in es.arg_to_shape.items():
                    ax = _long_axis(shape)
                    a = arrays[name]
                    if ax is None:
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
                for oname, odt in zip(es.output_names, out_dtypes):
                    cshape = tuple(
                        n if (out_axis is not None and k == out_axis) else out_shape[k]
                        for k in range(len(out_shape))
                    )
                    dev_out[oname] = self._dev_buf((slot, oname, n), cshape, odt)
                self._exec(self._run_q, **dev_in, **dev_out)
                ev_run[slot] = torch.cuda.Event()
                ev_run[slot].record(s_run)
                # ---- D2H
                s_d2h.wait_event(ev_run[slot])
                for oname in es.output_names:
                    o, buf = outs[oname], dev_out[oname]
                    if out_axis is None:
                        self._copy(o.ctypes.data, o.nbytes, buf.data_ptr(), o.nbytes, o.nbytes, 1, 1, s_d2h)
                        self.d2h_bytes += o.nbytes
                    else:
                        inner = int(np.prod(o.shape[out_axis + 1:], dtype=np.int64)) * o.itemsize
                        outer = int(np.prod(o.shape[:out_axis], dtype=np.int64))
                        self._copy(o.ctypes.data + lo * inner, o.shape[out_axis] * inner,
                                   buf.data_ptr(), n * inner, n * inner, outer, 1, s_d2h)
                        self.d2h_bytes += outer * n * inner
                ev_d2h[slot] = torch.cuda.Event()
                ev_d2h[slot].record(s_d2h)
            for s in (s_h2d, s_run, s_d2h):
                s.synchronize()
        return outs
