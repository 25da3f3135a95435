"""
Small helpers shared by the front-end (reference ``src/feinsum/utils.py:17-99``).
The TCCG benchmark table of the reference (``utils.py:103-233``) is generic
tensor-contraction material and out of scope for the DG hot path.
"""

from __future__ import annotations

import dataclasses as dc

from feinsum_b200.einsum import BatchedEinsum, SizeParam, SummationAxis


def is_any_redn_dim_parametric(einsum: BatchedEinsum) -> bool:
    """True if a contracted index has a symbolic (:class:`SizeParam`) extent."""
    descr = einsum.index_to_access_descr
    return any(
        isinstance(extent, SizeParam) and isinstance(descr[idx], SummationAxis)
        for idx, extent in einsum.index_to_dim_length.items()
    )


def get_n_redn_dim(einsum: BatchedEinsum) -> int:
    """Number of contracted indices."""
    return len(einsum.sum_indices)


@dc.dataclass
class IndexNameGenerator:
    """Hands out ``'a', 'b', ...`` skipping ``banned_names``; raises
    ``RuntimeError`` after ``'z'`` (reference ``utils.py:67-99``)."""

    banned_names: frozenset[str] = dc.field(default=frozenset())
    counter: int = dc.field(init=False, default=0)

    def __call__(self) -> str:
        while True:
            if self.counter >= 26:
                raise RuntimeError("All indices have been exhausted")
            name = chr(ord("a") + self.counter)
            self.counter += 1
            if name not in self.banned_names:
                return name
