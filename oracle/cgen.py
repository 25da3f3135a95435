"""
C/OpenMP restatement of the reference's code generator.  TEST INFRASTRUCTURE.

``generate_loopy(einsum, schedule)`` (reference
``src/feinsum/codegen/loopy.py:112-325``) emits, for every row of the batch
and every step of the contraction schedule, one statement

    result[out idx] = sum(reduction idx, prod_k operand_k[idx_k])

over a hyper-rectangular domain, with the symbolic axis a run-time integer,
inputs sorted by name, intermediates of a multi-step schedule kept in global
temporaries (``codegen/loopy.py:263-271``).  loopy lowers that to OpenCL C and
pocl runs it on the host cores -- none of which exists in this image.  This
module writes the same loop nest as plain C (symbolic axis outermost and
``omp parallel for``; ``-O3 -ffast-math -fopenmp`` standing in for the
reference's ``-cl-fast-relaxed-math -cl-mad-enable``, ``measure.py:133``),
compiles it with gcc and calls it through ctypes.

It is the "CPU restatement of the reference path (pocl unavailable)" that
``bench.py`` times, and a second, independent checker beside ``np_oracle``.
"""

from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from typing import Any

import numpy as np

_BUILD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")
_CTYPE = {np.dtype("float64"): "double", np.dtype("float32"): "float"}


def _is_int(x: Any) -> bool:
    return isinstance(x, (int, np.integer))


def _ext(e: Any) -> str:
    """C spelling of an extent: literal, or the symbolic parameter's name."""
    return str(int(e)) if _is_int(e) else str(getattr(e, "name", e))


def _split(subscripts: str) -> tuple[tuple[str, ...], list[tuple[str, ...]]]:
    ins, out = subscripts.replace(" ", "").split("->")
    return tuple(out), [tuple(s) for s in ins.split(",")]


def _flat_index(name_idx: tuple[str, ...], shape: tuple[Any, ...]) -> str:
    """Row-major linearisation with symbolic extents spelled as C variables."""
    if not name_idx:
        return "0"
    expr = name_idx[0]
    for idx, ext in zip(name_idx[1:], shape[1:]):
        expr = f"({expr})*{_ext(ext)} + {idx}"
    return expr


def generate_c(einsum: Any, schedule: Any | None = None) -> str:
    """C source of ``void fe_kernel(long long N, void **ins, void **outs)``.

    ``ins``: distinct operands sorted by name (reference
    ``codegen/loopy.py:146-152``); ``outs``: ``_fe_out``, ``_fe_out_0``, ...
    All symbolic extents are bound to ``N`` (``measure.py:135-139`` does the
    same with ``fix_parameters``).
    """
    if schedule is None:
        # trivial schedule, reference contraction_schedule.py:101-110
        steps = [(einsum.get_subscripts(), "_fe_out", tuple(range(einsum.n)))]
    else:
        steps = []
        for sub, res, operands in zip(
            schedule.subscripts, schedule.result_names, schedule.arguments
        ):
            ops = tuple(
                o.ioperand if hasattr(o, "ioperand") else o.name for o in operands
            )
            steps.append((sub, res, ops))

    params = sorted(p.name for p in einsum.all_size_params)
    in_names = sorted(einsum.all_args)
    lines = [
        "#include <stdlib.h>",
        "void fe_kernel(long long N_, void **ins, void **outs) {",
    ]
    for p in params:
        lines.append(f"  const long long {p} = N_;")

    for irow, row in enumerate(einsum.args):
        lines.append(f"  /* ---- row {irow} ---- */ {{")
        # name -> (C expression of base pointer, shape, dtype)
        avail: dict[Any, tuple[str, tuple[Any, ...], np.dtype]] = {}
        for k, arg in enumerate(row):
            cty = _CTYPE[np.dtype(arg.dtype)]
            avail[k] = (
                f"((const {cty}*)ins[{in_names.index(arg.name)}])",
                tuple(arg.shape),
                np.dtype(arg.dtype),
            )
        temporaries = []
        for istep, (sub, res, ops) in enumerate(steps):
            out_idx, in_idx_sets = _split(sub)
            extent: dict[str, Any] = {}
            for o, idx_set in zip(ops, in_idx_sets):
                for idx, ext in zip(idx_set, avail[o][1]):
                    extent[idx] = ext
            res_dtype = np.result_type(*[avail[o][2] for o in ops])
            cty = _CTYPE[np.dtype(res_dtype)]
            res_shape = tuple(extent[i] for i in out_idx)
            last = istep == len(steps) - 1
            if last:
                ptr = f"(({cty}*)outs[{irow}])"
            else:
                nelem = "*".join(f"(size_t){_ext(e)}" for e in res_shape) or "1"
                var = f"tmp_{irow}_{istep}"
                lines.append(
                    f"    {cty} *{var} = ({cty}*)malloc(sizeof({cty})*{nelem});"
                )
                temporaries.append(var)
                ptr = var
            avail[res] = (ptr, res_shape, np.dtype(res_dtype))

            def ub(idx: str) -> str:
                return _ext(extent[idx])

            long_free = [i for i in out_idx if not _is_int(extent[i])]
            short_free = [i for i in out_idx if _is_int(extent[i])]
            redn = sorted(
                {i for s in in_idx_sets for i in s} - set(out_idx)
            )
            order = long_free + short_free
            ind = "    "
            if order:
                lines.append(f"{ind}#pragma omp parallel for schedule(static)")
            for idx in order:
                lines.append(
                    f"{ind}for (long long {idx} = 0; {idx} < {ub(idx)}; ++{idx}) {{"
                )
                ind += "  "
            lines.append(f"{ind}{cty} acc = 0;")
            for idx in redn:
                lines.append(
                    f"{ind}for (long long {idx} = 0; {idx} < {ub(idx)}; ++{idx}) {{"
                )
                ind += "  "
            prod = " * ".join(
                f"{avail[o][0]}[{_flat_index(s, avail[o][1])}]"
                for o, s in zip(ops, in_idx_sets)
            )
            lines.append(f"{ind}acc += {prod};")
            for _ in redn:
                ind = ind[:-2]
                lines.append(f"{ind}}}")
            lines.append(f"{ind}{ptr}[{_flat_index(out_idx, res_shape)}] = acc;")
            for _ in order:
                ind = ind[:-2]
                lines.append(f"{ind}}}")
        for var in temporaries:
            lines.append(f"    free({var});")
        lines.append("  }")
    lines.append("}")
    return "\n".join(lines) + "\n"


def _cpu_signature() -> str:
    """-march=native objects are only valid on the CPU model that built them
    (build container and GPU box differ), so the cache key carries the flags line."""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown-cpu"


class CKernel:
    """gcc-compiled loop nest for one (einsum, schedule)."""

    def __init__(self, einsum: Any, schedule: Any | None = None, threads: bool = True):
        self.einsum = einsum
        self.in_names = sorted(einsum.all_args)
        self.source = generate_c(einsum, schedule)
        flags = ["-O3", "-ffast-math", "-march=native", "-fPIC", "-shared"]
        if threads:
            flags.append("-fopenmp")
        key = hashlib.sha1(
            (self.source + " ".join(flags) + _cpu_signature()).encode()
        ).hexdigest()[:16]
        os.makedirs(_BUILD_DIR, exist_ok=True)
        so = os.path.join(_BUILD_DIR, f"fe_{key}.so")
        if not os.path.exists(so):
            src = os.path.join(_BUILD_DIR, f"fe_{key}.c")
            with open(src, "w") as fh:
                fh.write(self.source)
            subprocess.run(["gcc", *flags, src, "-o", so + ".tmp"], check=True)
            os.replace(so + ".tmp", so)
        self._lib = ctypes.CDLL(so)
        self._fn = self._lib.fe_kernel
        self._fn.argtypes = [
            ctypes.c_longlong,
            ctypes.POINTER(ctypes.c_void_p),
            ctypes.POINTER(ctypes.c_void_p),
        ]
        self._fn.restype = None

    def out_shapes(self, n: int) -> list[tuple[int, ...]]:
        return [
            tuple(int(d) if _is_int(d) else n for d in self.einsum.shape)
            for _ in range(self.einsum.b)
        ]

    def __call__(
        self, n: int, arrays: dict[str, np.ndarray],
        outs: list[np.ndarray] | None = None,
    ) -> dict[str, np.ndarray]:
        ins = [np.ascontiguousarray(arrays[name]) for name in self.in_names]
        if outs is None:
            outs = []
            for row, shp in zip(self.einsum.args, self.out_shapes(n)):
                dt = np.result_type(*[a.dtype for a in row])
                outs.append(np.empty(shp, dtype=dt))
        in_ptrs = (ctypes.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
        out_ptrs = (ctypes.c_void_p * len(outs))(*[a.ctypes.data for a in outs])
        self._fn(int(n), in_ptrs, out_ptrs)
        names = ["_fe_out", *[f"_fe_out_{i}" for i in range(self.einsum.b - 1)]]
        return dict(zip(names, outs))
