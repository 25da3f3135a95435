"""
wave_3d_p4: the DG wave operator's three einsums -- div(v), grad(u) and the
four-field face-mass lift -- issued as ONE call on one stream.

In the reference this is a single loopy translation unit whose three statement
groups are separated by global barriers and transformed one by one
(reference ``examples/wave_3d_p4_auto.py:16-63,118-125``), i.e. three device
kernels behind one executor call.  Here it is
``fnsm_b200_wave3d_fused`` (``include/fnsm_b200.h``): the three DMMA kernels
back to back on the caller's stream, sharing ``J`` and the operator matrices.
(A single persistent kernel was rejected in round 1: the three operator tables
plus a slot and an output stage large enough for every item type leave shared
memory for 5-6 warps per SM, and these kernels need 10-12 to hide their
non-DMMA phases -- DESIGN.md section 4.)

Operand names and layouts follow the example: ``J(3,3,E) D(3,35,35) v(3,E,35)
u(E,35) L(35,4,15) Jface(4,E) F_0..F_3(4,E,15)`` ->
``div_out(E,35) grad_out(3,E,35) lift_0..lift_3(E,35)``.
"""

from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np

from feinsum_b200 import _cabi
from feinsum_b200.cl_utils import as_queue
from feinsum_b200.codegen.cuda import LaunchEvent
from feinsum_b200.einsum import BatchedEinsum
from feinsum_b200.make_einsum import array, batched_einsum, einsum

INPUTS = ("J", "D", "v", "u", "L", "Jface", "F_0", "F_1", "F_2", "F_3")
OUTPUTS = ("div_out", "grad_out", "lift_0", "lift_1", "lift_2", "lift_3")


def wave3d_einsums(dtype: Any = "float64") -> dict[str, BatchedEinsum]:
    """The three einsums of the operator (same arrays as the fused call)."""
    J, D = array("J", (3, 3, "E"), dtype), array("D", (3, 35, 35), dtype)
    return {
        "div": einsum("xre,rij,xej->ei", J, D, array("v", (3, "E", 35), dtype)),
        "grad": einsum("xre,rij,ej->xei", J, D, array("u", ("E", 35), dtype)),
        "lift": batched_einsum(
            "ifj,fe,fej->ei",
            [[array("L", (35, 4, 15), dtype), array("Jface", (4, "E"), dtype),
              array(f"F_{k}", (4, "E", 15), dtype)] for k in range(4)]),
    }


def shapes(n_elements: int) -> tuple[dict[str, tuple[int, ...]], dict[str, tuple[int, ...]]]:
    E = int(n_elements)
    ins = {"J": (3, 3, E), "D": (3, 35, 35), "v": (3, E, 35), "u": (E, 35), "L": (35, 4, 15),
           "Jface": (4, E), **{f"F_{k}": (4, E, 15) for k in range(4)}}
    outs = {"div_out": (E, 35), "grad_out": (3, E, 35), **{f"lift_{k}": (E, 35) for k in range(4)}}
    return ins, outs


#: work model of the fused operator per element: J is read once
FLOPS_PER_ELEMENT = 7980 + 7980 + 17040
BYTES_PER_ELEMENT = {np.dtype("float64"): 5384, np.dtype("float32"): 2692}


class Wave3DProgram:
    """The operator as a *program* (what ``generate_cuda`` returns for a single einsum): yields the
    device executor, and describes its host boundary so that
    :class:`feinsum_b200.host_exec.HostExecutor` can pipeline it over element chunks."""

    kernel_id = "wave3d"

    def __init__(self, dtype: Any = "float64", **params: int):
        self.dtype = np.dtype(dtype)
        self.params = dict(params)

    def with_params(self, **params: int) -> "Wave3DProgram":
        return Wave3DProgram(self.dtype, **{**self.params, **params})

    def executor(self, cq: Any = None, **_unused: Any) -> "Wave3DExecutor":
        return Wave3DExecutor(cq, self.dtype, **self.params)

    def host_spec(self) -> Any:
        from feinsum_b200.einsum import SizeParam
        from feinsum_b200.host_exec import HostSpec

        E = SizeParam("E")
        ins, outs = shapes(0)
        sym = lambda shp, pos: tuple(E if k == pos else d for k, d in enumerate(shp))  # noqa: E731
        e_axis = {"J": 2, "v": 1, "u": 0, "Jface": 1, **{f"F_{k}": 1 for k in range(4)},
                  "div_out": 0, "grad_out": 1, **{f"lift_{k}": 0 for k in range(4)}}
        return HostSpec(
            {n: (sym(s, e_axis[n]) if n in e_axis else s) for n, s in ins.items()},
            {n: self.dtype for n in ins},
            {n: sym(s, e_axis[n]) for n, s in outs.items()},
            {n: self.dtype for n in outs},
        )


class Wave3DExecutor:
    """``evt, outs = Wave3DExecutor(cq)(cq, J=..., D=..., ...)`` -- torch CUDA tensors in,
    dict of the six outputs back (pre-allocated outputs may be passed by name)."""

    def __init__(self, cq: Any = None, dtype: Any = "float64", **params: int):
        self.cq = as_queue(cq)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype("float64"), np.dtype("float32")):
            raise NotImplementedError("wave3d: float64 or float32")
        self.lib = _cabi.lib()
        self._cfg = _cabi.make_cfg(params)

    def __call__(self, cq: Any = None, allocator: Any = None, **arrays: Any) -> tuple[LaunchEvent, dict[str, Any]]:
        import torch

        q = self.cq if cq is None else as_queue(cq)
        tdt = torch.float64 if self.dtype == np.dtype("float64") else torch.float32
        missing = [n for n in INPUTS if n not in arrays]
        if missing:
            raise TypeError(f"missing input arrays {missing}")
        unknown = set(arrays) - set(INPUTS) - set(OUTPUTS)
        if unknown:
            raise TypeError(f"unexpected arguments: {sorted(unknown)}")
        E = int(arrays["u"].shape[0])
        in_shapes, out_shapes = shapes(E)
        for n in INPUTS:
            a = arrays[n]
            if not isinstance(a, torch.Tensor) or not a.is_cuda or a.device.index != q.device.index:
                raise TypeError(f"'{n}' must be a torch CUDA tensor on {q.torch_device}")
            if tuple(a.shape) != in_shapes[n] or a.dtype != tdt or not a.is_contiguous():
                raise ValueError(f"'{n}': expected C-contiguous {self.dtype} of shape {in_shapes[n]}, "
                                 f"got {a.dtype} {tuple(a.shape)}")
        outs: dict[str, Any] = {}
        for n in OUTPUTS:
            o = arrays.get(n)
            if o is None:
                o = torch.empty(out_shapes[n], dtype=tdt, device=q.torch_device)
            elif tuple(o.shape) != out_shapes[n] or o.dtype != tdt or not o.is_contiguous():
                raise ValueError(f"output '{n}' has wrong shape/dtype/layout")
            outs[n] = o
        wa = _cabi.WaveArgs()
        for n in ("J", "D", "v", "u", "L", "Jface"):
            setattr(wa, n, arrays[n].data_ptr())
        for k in range(4):
            wa.F[k] = arrays[f"F_{k}"].data_ptr()
            wa.lift_out[k] = outs[f"lift_{k}"].data_ptr()
        wa.div_out = outs["div_out"].data_ptr()
        wa.grad_out = outs["grad_out"].data_ptr()
        with torch.cuda.device(q.torch_device), torch.cuda.stream(q.torch_stream):
            if E > 0:
                rc = self.lib.fnsm_b200_wave3d_fused(
                    _cabi.FNSM_F64 if self.dtype == np.dtype("float64") else _cabi.FNSM_F32,
                    C.byref(wa), C.c_int64(E), self._cfg, C.c_void_p(q.stream))
                _cabi.check(rc, "fnsm_b200_wave3d_fused")
            evt = torch.cuda.Event()
            evt.record(q.torch_stream)
        return LaunchEvent(evt), outs
