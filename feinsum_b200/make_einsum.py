"""
User-facing constructors: :func:`array`, :func:`einsum`, :func:`batched_einsum`.

Behavioural restatement of the reference's ``feinsum.make_einsum``
(reference ``src/feinsum/make_einsum.py:55-156``): explicit-mode subscripts
only, one ASCII letter per index, blanks ignored, ``...`` rejected with
``NotImplementedError``, repeated output index rejected with ``ValueError``,
every consistency failure of :class:`BatchedEinsum` surfaced as ``TypeError``.
"""

from __future__ import annotations

from collections.abc import Iterable, Sequence
from typing import Any

import numpy as np
import numpy.typing as npt

from feinsum_b200.einsum import (
    INT_CLASSES,
    Array,
    BatchedEinsum,
    ShapeComponentT,
    ShapeT,
    SizeParam,
)


def _shape_component(dim: Any) -> ShapeComponentT:
    # reference make_einsum.py:55-61
    if isinstance(dim, str):
        return SizeParam(dim)
    if isinstance(dim, SizeParam):
        return dim
    if isinstance(dim, INT_CLASSES) and dim >= 0:
        return dim
    raise ValueError(f"Cannot infer shape component '{dim}'.")


def _shape(shape: Any) -> ShapeT:
    # reference make_einsum.py:64-70 (a bare ``str`` is iterated per
    # character there as well; callers pass tuples or ints)
    if not isinstance(shape, Iterable):
        shape = (shape,)
    return tuple(_shape_component(dim) for dim in shape)


def array(name: str, shape: Any, dtype: npt.DTypeLike = "float64") -> Array:
    """``array("J", (3, 3, "E"))`` -- a ``str`` extent is a :class:`SizeParam`.

    reference: ``src/feinsum/make_einsum.py:73-77``.
    """
    return Array(name=name, shape=_shape(shape), dtype=np.dtype(dtype))


class _BroadcastingNotSupported(NotImplementedError, TypeError):
    """``...`` in a subscript.  The reference means to raise
    ``NotImplementedError("Broadcasting in einsums not supported")``
    (``make_einsum.py:98``) but, because ``groupdict()`` always holds the
    ``alpha`` key, it actually appends ``None`` and dies with a ``TypeError``
    in ``BatchedEinsum.__post_init__`` (pinned in ``tests/golden/frontend.json``).
    This type satisfies handlers written against either behaviour."""


def _parse_indices(spec: str, *, is_output: bool) -> tuple[str, ...]:
    """Split one operand's subscript into index letters.

    reference: ``src/feinsum/make_einsum.py:83-111``.
    """
    indices: list[str] = []
    pos, end = 0, len(spec)
    while pos < end:
        ch = spec[pos]
        if ch.isspace():
            pos += 1
        elif spec.startswith("...", pos):
            raise _BroadcastingNotSupported("Broadcasting in einsums not supported")
        elif ch.isascii() and ch.isalpha():
            indices.append(ch)
            pos += 1
        else:
            raise ValueError(
                f"Cannot parse '{spec[pos:]}' in provided einsum '{spec}'."
            )
    if is_output and len(set(indices)) != len(indices):
        raise ValueError(
            f"Used an input more than once to refer to the output axis in '{spec}"
        )
    return tuple(indices)


def parse_subscripts(
    subscripts: str,
) -> tuple[tuple[str, ...], tuple[tuple[str, ...], ...]]:
    """``"xre,rij,ej->xei"`` -> ``(out_idx_set, in_idx_sets)``."""
    if "->" not in subscripts:
        raise ValueError(
            "Missing -> in 'subscripts'. If the expected behavior"
            " is implicit mode, feinsum does not support it."
        )
    in_specs, out_spec = subscripts.split("->")
    out_idx_set = _parse_indices(out_spec, is_output=True)
    in_idx_sets = tuple(
        _parse_indices(spec, is_output=False) for spec in in_specs.split(",")
    )
    return out_idx_set, in_idx_sets


def batched_einsum(
    subscripts: str, args: Sequence[Sequence[Array]]
) -> BatchedEinsum:
    """``numpy.einsum``-like constructor for ``b`` rows of ``n`` operands.

    reference: ``src/feinsum/make_einsum.py:114-148``.
    """
    out_idx_set, in_idx_sets = parse_subscripts(subscripts)
    try:
        return BatchedEinsum(
            out_idx_set, in_idx_sets, tuple(tuple(row) for row in args)
        )
    except AssertionError as exc:
        raise TypeError(f"{exc}") from exc


def einsum(subscripts: str, *operands: Array) -> BatchedEinsum:
    """One-row :func:`batched_einsum` (reference ``make_einsum.py:151-156``)."""
    return batched_einsum(subscripts, [operands])
