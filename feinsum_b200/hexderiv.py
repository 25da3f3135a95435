"""
Fused hex derivative: the three tensor-product (sum-factorisation) modes of a p = 7 hexahedral element applied to
the SAME field in ONE call --

    d0 = eabc,ia->eibc (M0)     d1 = eabc,ib->eaic (M1)     d2 = eabc,ic->eabi (M2)

``A(E,8,8,8)`` is read once: 16 KB of traffic per fp64 element instead of the 24 KB of three separate
``eabc,ia->eibc``-style einsums (each of which is ``fnsm_b200_tensor_product``).  The three einsums have different
subscripts, so -- like the wave operator (:mod:`feinsum_b200.wave3d`) -- the fused form is a *program* of several
einsums behind one executor call, not a single ``BatchedEinsum``; :func:`hexderiv_einsums` returns the three
einsums for the oracle / the CPU restatement.  The reference has no transform for tensor-product einsums at all
(SURVEY.md section 2): each would run ``generate_loopy``'s trivial schedule.

C ABI: ``fnsm_b200_hex_deriv`` (``include/fnsm_b200.h``), kernel ``csrc/hex_deriv.cu``.
"""

from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np

from feinsum_b200 import _cabi
from feinsum_b200.cl_utils import as_queue
from feinsum_b200.codegen.cuda import LaunchEvent
from feinsum_b200.einsum import BatchedEinsum
from feinsum_b200.make_einsum import array, einsum

N1D = 8
INPUTS = ("A", "M0", "M1", "M2")
OUTPUTS = ("d0", "d1", "d2")
SUBSCRIPTS = ("eabc,ia->eibc", "eabc,ib->eaic", "eabc,ic->eabi")

#: work model per element: three applications of an 8 x 8 operator to 512 values; A once, three outputs
FLOPS_PER_ELEMENT = 3 * 2 * N1D * N1D**3
BYTES_PER_ELEMENT = {np.dtype("float64"): 4 * N1D**3 * 8, np.dtype("float32"): 4 * N1D**3 * 4}


def hexderiv_einsums(dtype: Any = "float64") -> dict[str, BatchedEinsum]:
    A = array("A", ("E", N1D, N1D, N1D), dtype)
    return {f"d{k}": einsum(sub, A, array(f"M{k}", (N1D, N1D), dtype)) for k, sub in enumerate(SUBSCRIPTS)}


def shapes(n_elements: int) -> tuple[dict[str, tuple[int, ...]], dict[str, tuple[int, ...]]]:
    E = int(n_elements)
    full = (E, N1D, N1D, N1D)
    return ({"A": full, **{f"M{k}": (N1D, N1D) for k in range(3)}}, {f"d{k}": full for k in range(3)})


class HexDerivProgram:
    """What ``generate_cuda`` is for a single einsum: yields the executor and describes the host boundary
    (:class:`feinsum_b200.host_exec.HostExecutor` pipelines it over element chunks)."""

    kernel_id = "hex_deriv"

    def __init__(self, dtype: Any = "float64", **params: int):
        self.dtype = np.dtype(dtype)
        self.params = dict(params)

    def with_params(self, **params: int) -> "HexDerivProgram":
        return HexDerivProgram(self.dtype, **{**self.params, **params})

    def executor(self, cq: Any = None, **_unused: Any) -> "HexDerivExecutor":
        return HexDerivExecutor(cq, self.dtype, **self.params)

    def host_spec(self) -> Any:
        from feinsum_b200.einsum import SizeParam
        from feinsum_b200.host_exec import HostSpec

        full = (SizeParam("E"), N1D, N1D, N1D)
        return HostSpec({"A": full, **{f"M{k}": (N1D, N1D) for k in range(3)}}, {n: self.dtype for n in INPUTS},
                        {n: full for n in OUTPUTS}, {n: self.dtype for n in OUTPUTS})


class HexDerivExecutor:
    """``evt, outs = HexDerivExecutor(cq)(cq, A=..., M0=..., M1=..., M2=...)`` -- torch CUDA tensors in, dict of
    ``d0, d1, d2`` back (pre-allocated outputs may be passed by name)."""

    def __init__(self, cq: Any = None, dtype: Any = "float64", **params: int):
        self.cq = as_queue(cq)
        self.dtype = np.dtype(dtype)
        if self.dtype != np.dtype("float64"):
            raise NotImplementedError("fused hex derivative: float64 (use three tensor-product einsums for float32)")
        self.lib = _cabi.lib()
        self._cfg = _cabi.make_cfg(params)

    def __call__(self, cq: Any = None, allocator: Any = None, **arrays: Any) -> tuple[LaunchEvent, dict[str, Any]]:
        import torch

        q = self.cq if cq is None else as_queue(cq)
        missing = [n for n in INPUTS if n not in arrays]
        if missing:
            raise TypeError(f"missing input arrays {missing}")
        unknown = set(arrays) - set(INPUTS) - set(OUTPUTS)
        if unknown:
            raise TypeError(f"unexpected arguments: {sorted(unknown)}")
        E = int(arrays["A"].shape[0])
        in_shapes, out_shapes = shapes(E)
        for n in INPUTS:
            a = arrays[n]
            if not isinstance(a, torch.Tensor) or not a.is_cuda or a.device.index != q.device.index:
                raise TypeError(f"'{n}' must be a torch CUDA tensor on {q.torch_device}")
            if tuple(a.shape) != in_shapes[n] or a.dtype != torch.float64 or not a.is_contiguous():
                raise ValueError(f"'{n}': expected C-contiguous float64 of shape {in_shapes[n]}, "
                                 f"got {a.dtype} {tuple(a.shape)}")
        outs: dict[str, Any] = {}
        for n in OUTPUTS:
            o = arrays.get(n)
            if o is None:
                o = torch.empty(out_shapes[n], dtype=torch.float64, device=q.torch_device)
            elif tuple(o.shape) != out_shapes[n] or o.dtype != torch.float64 or not o.is_contiguous():
                raise ValueError(f"output '{n}' has wrong shape/dtype/layout")
            outs[n] = o
        mats = (C.c_void_p * 3)(*[arrays[f"M{k}"].data_ptr() for k in range(3)])
        optr = (C.c_void_p * 3)(*[outs[f"d{k}"].data_ptr() for k in range(3)])
        with torch.cuda.device(q.torch_device), torch.cuda.stream(q.torch_stream):
            if E > 0:
                rc = self.lib.fnsm_b200_hex_deriv(_cabi.FNSM_F64, C.c_void_p(arrays["A"].data_ptr()), mats, optr,
                                                  N1D, C.c_int64(E), self._cfg, C.c_void_p(q.stream))
                _cabi.check(rc, "fnsm_b200_hex_deriv")
            evt = torch.cuda.Event()
            evt.record(q.torch_stream)
        return LaunchEvent(evt), outs
