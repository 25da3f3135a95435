"""
Small helpers shared by the front-end (reference ``src/feinsum/utils.py:17-99``).
plus the TCCG benchmark getter (``utils.py:103-233``): generic tensor contractions that
exercise the generic CUDA kernel.
"""

from __future__ import annotations

import dataclasses as dc
from typing import Any

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


# ---------------------------------------------------------------------------
# TCCG benchmark suite (Springer & Bientinesi; the input strings of the COGENT
# artifact, as tabulated in the reference's ``utils.py:103-233``).  One line per
# contraction: ``<output>-<A>-<B> <extent of a>,<extent of b>,...``.
_TCCG = """
    abc-bda-dc 312,312,24,312
    abc-dca-bd 312,24,296,312
    abcd-dbea-ec 72,72,24,72,72
    abcd-deca-be 72,24,72,72,72
    abcd-ebad-ce 72,72,24,72,72
    abcde-efbad-cf 48,32,24,32,48,32
    abcde-ecbfa-fd 48,32,32,24,48,48
    abcde-efcad-bf 48,24,32,32,48,32
    abcd-ea-ebcd 72,72,72,72,72
    abcd-eb-aecd 72,72,72,72,72
    abcd-ec-abed 72,72,72,72,72
    ab-ac-cb 5136,5120,5136
    ab-acd-dbc 312,296,296,312
    ab-cad-dcb 312,296,312,312
    abc-acd-db 312,296,296,312
    abc-ad-bdc 312,312,296,296
    abc-adc-bd 312,312,296,296
    abc-adc-db 312,296,296,312
    abc-adec-ebd 72,72,72,72,72
    abcd-aebf-dfce 72,72,72,72,72,72
    abcd-aebf-fdec 72,72,72,72,72,72
    abcd-aecf-bfde 72,72,72,72,72,72
    abcd-aecf-fbed 72,72,72,72,72,72
    abcd-aedf-bfce 72,72,72,72,72,72
    abcd-aedf-fbec 72,72,72,72,72,72
    abcd-aefb-fdce 72,72,72,72,72,72
    abcd-aefc-fbed 72,72,72,72,72,72
    abcd-eafb-fdec 72,72,72,72,72,72
    abcd-eafc-bfde 72,72,72,72,72,72
    abcd-eafd-fbec 72,72,72,72,72,72
    abcdef-dega-gfbc 24,16,16,24,16,16,24
    abcdef-degb-gfac 24,16,16,24,16,16,24
    abcdef-degc-gfab 24,16,16,24,16,16,24
    abcdef-dfga-gebc 24,16,16,24,16,16,24
    abcdef-dfgb-geac 24,16,16,24,16,16,24
    abcdef-dfgc-geab 24,16,16,24,16,16,24
    abcdef-efga-gdbc 24,16,16,16,24,16,24
    abcdef-efgb-gdac 24,16,16,16,24,16,24
    abcdef-efgc-gdab 24,16,16,16,24,16,24
    abcdef-gdab-efgc 24,16,16,16,24,16,24
    abcdef-gdac-efgb 24,16,16,16,24,16,24
    abcdef-gdbc-efga 24,16,16,16,24,16,24
    abcdef-geab-dfgc 24,16,16,24,16,16,24
    abcdef-geac-dfgb 24,16,16,24,16,16,24
    abcdef-gebc-dfga 24,16,16,24,16,16,24
    abcdef-gfab-degc 24,16,16,24,16,16,24
    abcdef-gfac-degb 24,16,16,24,16,16,24
    abcdef-gfbc-dega 24,16,16,24,16,16,24
""".split("\n")[1:-1]


def get_tccg_benchmark(i: int, dtype: Any = "float64") -> BatchedEinsum:
    """The *i*-th (1-based, 48 in all) tensor contraction of the TCCG suite as a
    two-operand :class:`BatchedEinsum` with operands ``A`` and ``B`` (reference
    ``utils.get_tccg_benchmark``, ``src/feinsum/utils.py:206-233``).  They run on the
    generic CUDA kernel (``classify`` -> ``generic``)."""
    from feinsum_b200.make_einsum import array, einsum

    if not (isinstance(i, int) and 1 <= i <= len(_TCCG)):
        raise ValueError(f"i must be in the set {{1, 2, .., {len(_TCCG)}}}. Got {i = }.")
    spec, extents = _TCCG[i - 1].split()
    out, in_a, in_b = spec.split("-")
    length = {chr(ord("a") + k): int(n) for k, n in enumerate(extents.split(","))}
    return einsum(
        f"{in_a},{in_b}->{out}",
        array("A", [length[c] for c in in_a], dtype),
        array("B", [length[c] for c in in_b], dtype),
    )
