"""
Batched-einsum description objects (front-end IR).

API-compatible restatement of the reference's ``feinsum.einsum`` module
(reference ``src/feinsum/einsum.py:26-387``): ``SizeParam``, ``Array``,
``EinsumAxisAccess``/``FreeAxis``/``SummationAxis`` and ``BatchedEinsum``
with the same field names, validation order, derived properties and
``get_subscripts()`` spelling (``"xre,rij,ej -> xei"``).  It depends on
numpy only; ``immutables``/``pytools``/``islpy`` are not needed.
"""

from __future__ import annotations

from dataclasses import dataclass, replace
from functools import cached_property
from typing import Any

import numpy as np

from feinsum_b200._immutable import Map

IntegralT = int | np.integer
INT_CLASSES = (int, np.integer)


@dataclass(frozen=True)
class SizeParam:
    """A symbolic ("very long") axis length, e.g. the element count ``E``.

    reference: ``src/feinsum/einsum.py:26-41``.
    """

    name: str

    def __truediv__(self, other: Any) -> Any:
        # tuner parameter-getters may write ``shape[k] / 4``; like the
        # reference this is undefined for symbolic extents.
        return NotImplemented

    __rtruediv__ = __truediv__


ShapeComponentT = IntegralT | SizeParam
ShapeT = tuple[ShapeComponentT, ...]


@dataclass(frozen=True, eq=True, repr=True)
class Array:
    """A named n-d operand: ``name``, ``shape`` (ints / :class:`SizeParam`), ``dtype``.

    reference: ``src/feinsum/einsum.py:48-83``.
    """

    name: str
    shape: ShapeT
    dtype: np.dtype[Any]

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def copy(
        self,
        *,
        name: str | None = None,
        shape: ShapeT | None = None,
        dtype: np.dtype[Any] | None = None,
    ) -> "Array":
        changes: dict[str, Any] = {}
        if name is not None:
            changes["name"] = name
        if shape is not None:
            changes["shape"] = shape
        if dtype is not None:
            changes["dtype"] = dtype
        return replace(self, **changes)


@dataclass(frozen=True)
class EinsumAxisAccess:
    """Abstract: how an index of the einsum is used (free vs. summed)."""

    def __init__(self) -> None:
        if type(self) is EinsumAxisAccess:
            raise TypeError(
                "EinsumAxisAccess is abstract and cannot be instantiated directly"
            )


@dataclass(frozen=True)
class FreeAxis(EinsumAxisAccess):
    """Index that survives into the output at position ``output_index``."""

    output_index: int


@dataclass(frozen=True)
class SummationAxis(EinsumAxisAccess):
    """Contracted index; ``index`` numbers them in order of first appearance."""

    index: int


def _is_index_name(idx: Any) -> bool:
    return isinstance(idx, str) and len(idx) == 1 and idx.islower()


@dataclass(frozen=True)
class BatchedEinsum:
    """
    ``b`` einsums sharing one subscript expression, each on its own row of
    ``n`` operands.  Output ``k`` is conventionally called ``_fe_out`` (k = 0)
    or ``_fe_out_{k-1}``.

    reference: ``src/feinsum/einsum.py:127-387`` (checks in ``:159-196``).
    """

    out_idx_set: tuple[str, ...]
    in_idx_sets: tuple[tuple[str, ...], ...]
    args: tuple[tuple[Array, ...], ...]

    def __post_init__(self) -> None:
        # Same checks, same order and same messages as the reference so that
        # ``batched_einsum`` re-raises identical ``TypeError`` texts.
        assert all(
            _is_index_name(idx) for idx in self.out_idx_set
        ), "Obtained invalid output index (RHS of ->)."
        assert all(
            _is_index_name(idx) for idx_set in self.in_idx_sets for idx in idx_set
        ), "Obtained invalid input index (LHS of ->)."

        seen_in: set[str] = set()
        for idx_set in self.in_idx_sets:
            seen_in.update(idx_set)
        assert (
            set(self.out_idx_set) <= seen_in
        ), "Obtained an out index which is not present in the input indices."

        n_operands = len(self.in_idx_sets)
        assert all(
            len(row) == n_operands for row in self.args
        ), "Mismatch in #operands between subscript expression and input arrays."
        assert all(
            arg.ndim == len(idx_set)
            for row in self.args
            for arg, idx_set in zip(row, self.in_idx_sets)
        ), "Dimensionality of input operands do no match the provided subscripts."

        # force evaluation: these raise AssertionError on inconsistencies
        _ = self.arg_to_dtype
        _ = self.arg_to_shape
        _ = self.index_to_dim_length

        param_names = {p.name for p in self.all_size_params}
        n_names = len(self.all_args) + len(self.all_indices) + len(param_names)
        assert n_names == len(
            set(self.all_args) | set(self.all_indices) | param_names
        ), "Must use different names for arguments, indices, and size params."

    # -- sizes -----------------------------------------------------------
    @cached_property
    def b(self) -> int:
        """Number of einsums (rows) in the batch."""
        return len(self.args)

    @cached_property
    def n(self) -> int:
        """Number of operands of every einsum of the batch."""
        return len(self.in_idx_sets)

    @cached_property
    def index_to_dim_length(self) -> Map[str, ShapeComponentT]:
        lengths: dict[str, ShapeComponentT] = {}
        for row in self.args:
            for arg, idx_set in zip(row, self.in_idx_sets):
                for idx, extent in zip(idx_set, arg.shape):
                    if idx not in lengths:
                        lengths[idx] = extent
                    elif lengths[idx] != extent:
                        raise AssertionError(
                            "Shape mismatch for indices across the arguments."
                        )
        return Map(lengths)

    @cached_property
    def shape(self) -> ShapeT:
        """Shape of each output."""
        return tuple(self.index_to_dim_length[idx] for idx in self.out_idx_set)

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def get_subscripts(self) -> str:
        """``"xre,rij,ej -> xei"`` (blanks around the arrow, as the reference)."""
        lhs = ",".join("".join(idx_set) for idx_set in self.in_idx_sets)
        return f"{lhs} -> {''.join(self.out_idx_set)}"

    @cached_property
    def arg_to_shape(self) -> Map[str, ShapeT]:
        shapes: dict[str, ShapeT] = {}
        for row in self.args:
            for arg in row:
                if shapes.setdefault(arg.name, arg.shape) != arg.shape:
                    raise AssertionError(f"Inconsistent shapes for arg {arg.name}.")
        return Map(shapes)

    @cached_property
    def arg_to_dtype(self) -> Map[str, np.dtype[Any]]:
        dtypes: dict[str, np.dtype[Any]] = {}
        for row in self.args:
            for arg in row:
                if dtypes.setdefault(arg.name, arg.dtype) != arg.dtype:
                    raise AssertionError(f"Inconsistent dtypes for arg {arg.name}.")
        return Map(dtypes)

    @cached_property
    def index_to_access_descr(self) -> Map[str, EinsumAxisAccess]:
        descr: dict[str, EinsumAxisAccess] = {
            idx: FreeAxis(pos) for pos, idx in enumerate(self.out_idx_set)
        }
        n_sum = 0
        for idx_set in self.in_idx_sets:
            for idx in idx_set:
                if idx not in descr:
                    descr[idx] = SummationAxis(n_sum)
                    n_sum += 1
        return Map(descr)

    @cached_property
    def sum_indices(self) -> tuple[str, ...]:
        """Contracted indices in order of first appearance."""
        numbered = [
            (acc.index, idx)
            for idx, acc in self.index_to_access_descr.items()
            if isinstance(acc, SummationAxis)
        ]
        return tuple(idx for _, idx in sorted(numbered))

    @cached_property
    def all_args(self) -> frozenset[str]:
        return frozenset(self.arg_to_shape)

    @cached_property
    def all_indices(self) -> frozenset[str]:
        return frozenset(self.index_to_dim_length)

    @cached_property
    def all_size_params(self) -> frozenset[SizeParam]:
        return frozenset(
            v for v in self.index_to_dim_length.values() if isinstance(v, SizeParam)
        )

    @property
    def output_names(self) -> tuple[str, ...]:
        """``("_fe_out", "_fe_out_0", ...)`` -- reference ``einsum.py:359``,
        ``measure.py:147``."""
        return ("_fe_out", *(f"_fe_out_{k}" for k in range(self.b - 1)))

    def copy(
        self,
        *,
        out_idx_set: tuple[str, ...] | None = None,
        in_idx_sets: tuple[tuple[str, ...], ...] | None = None,
        args: tuple[tuple[Array, ...], ...] | None = None,
    ) -> "BatchedEinsum":
        return BatchedEinsum(
            self.out_idx_set if out_idx_set is None else out_idx_set,
            self.in_idx_sets if in_idx_sets is None else in_idx_sets,
            self.args if args is None else args,
        )

    def _domain_str(self) -> str:
        # The reference pretty-prints an ISL set here (einsum.py:381); the
        # box constraints are written out by hand in the same notation.
        names = sorted(self.index_to_dim_length)
        params = sorted(p.name for p in self.all_size_params)
        cons = []
        for idx in names:
            ext = self.index_to_dim_length[idx]
            ub = ext.name if isinstance(ext, SizeParam) else int(ext)
            cons.append(f"0 <= {idx} < {ub}")
        prefix = f"[{', '.join(params)}] -> " if params else ""
        return f"{prefix}{{ [{', '.join(names)}] : {' and '.join(cons)} }}"

    def __str__(self) -> str:
        from tabulate import tabulate

        dtypes = "\n".join(
            f"{name}: {dtype}" for name, dtype in sorted(self.arg_to_dtype.items())
        )
        sum_idxs = "{" + ", ".join(self.sum_indices) + "}"
        out_idxs = ", ".join(self.out_idx_set)
        rows = []
        for out_name, row in zip(self.output_names, self.args):
            product = "×".join(  # noqa: RUF001
                f"{arg.name}[{', '.join(idx_set)}]"
                for arg, idx_set in zip(row, self.in_idx_sets)
            )
            rows.append([" ", f"{out_name}[{out_idxs}]", "<-", f"Σ_{sum_idxs} {product}"])
        statements = tabulate(
            rows, tablefmt="plain", colalign=("left", "right", "left", "left")
        )
        rule = "-" * 75
        return (
            f"{rule}\nDOMAINS:\n{self._domain_str()}\n{rule}\nData-types:\n{dtypes}\n"
            f"{rule}\nfor {','.join(self.out_idx_set)}\n{statements}\nend\n{rule}"
        )
