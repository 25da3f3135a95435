"""
Canonical form of a batched einsum (database key).

Two batched einsums are *isomorphic* if one is obtained from the other by
renaming indices, operands and size parameters, permuting the operands of the
subscript expression and permuting the rows of the batch; axis order inside an
operand and inside the output is significant (it is the memory layout).  This
is the notion the reference uses for its transform database (reference
``src/feinsum/canonicalization.py:1087-1125``, doc/design.rst), where it is
computed by building a coloured graph and handing it to the C++ ``pybliss``
canonical-labelling library -- which is not available here.

This module computes a canonical representative directly:

1. operand positions are ordered by a renaming-invariant key (rank, access
   pattern against the output, extents, dtype/sharing multiset of the column);
   only positions with equal keys are permuted exhaustively;
2. for a candidate operand order, indices are named ``a, b, ...`` by first
   appearance (output first, then operands left to right), size parameters
   after the first index that carries them;
3. rows and operand names are canonicalised by colour refinement on the
   bipartite (array, (row, position)) incidence structure, remaining ties broken
   by exhaustive search when small;
4. the lexicographically smallest signature over all candidates wins.

Canonical names: indices ``a..z``, arrays ``arg_0, arg_1, ...`` (first
appearance, row-major), size parameter of index ``d`` is ``D`` (the spelling the
reference's database layer reconstructs, ``sql_utils.py:316-320``).
"""

from __future__ import annotations

from itertools import permutations, product
from typing import Any

import numpy as np

from feinsum_b200._immutable import Map
from feinsum_b200.einsum import Array, BatchedEinsum, SizeParam

_MAX_TIE_PERMS = 5040


def _extent_key(ext: Any) -> tuple[int, int]:
    return (1, 0) if isinstance(ext, SizeParam) else (0, int(ext))


def _position_key(einsum: BatchedEinsum, k: int) -> tuple[Any, ...]:
    """Renaming-invariant description of operand position *k*."""
    idx_set = einsum.in_idx_sets[k]
    out = einsum.out_idx_set
    axes = []
    for pos, idx in enumerate(idx_set):
        first = idx_set.index(idx)
        axes.append(
            (
                _extent_key(einsum.index_to_dim_length[idx]),
                out.index(idx) if idx in out else -1,
                first if first != pos else -1,  # repeated-index pattern
                sum(1 for s in einsum.in_idx_sets for i in s if i == idx),
            )
        )
    column = sorted(
        (np.dtype(row[k].dtype).name, sum(1 for r in einsum.args if r[k].name == row[k].name))
        for row in einsum.args
    )
    return (len(idx_set), tuple(axes), tuple(column))


def _candidate_orders(einsum: BatchedEinsum) -> list[tuple[int, ...]]:
    keys = [_position_key(einsum, k) for k in range(einsum.n)]
    order = sorted(range(einsum.n), key=lambda k: keys[k])
    groups: list[list[int]] = []
    for k in order:
        if groups and keys[groups[-1][0]] == keys[k]:
            groups[-1].append(k)
        else:
            groups.append([k])
    total = 1
    for g in groups:
        for m in range(2, len(g) + 1):
            total *= m
    if total > _MAX_TIE_PERMS:
        return [tuple(order)]
    return [
        tuple(k for part in combo for k in part)
        for combo in product(*[list(permutations(g)) for g in groups])
    ]


def _refine_rows(
    einsum: BatchedEinsum, order: tuple[int, ...]
) -> list[list[int]]:
    """Colour refinement; returns row indices grouped by (sorted) colour class."""
    rows = [[row[k].name for k in order] for row in einsum.args]
    shape_sig = {
        name: tuple(_extent_key(d) for d in shape)
        for name, shape in einsum.arg_to_shape.items()
    }
    arg_col: dict[str, Any] = {
        name: (np.dtype(einsum.arg_to_dtype[name]).name, shape_sig[name])
        for name in einsum.all_args
    }
    row_col: list[Any] = [None] * len(rows)

    def compress(values: dict[Any, Any] | list[Any]) -> Any:
        items = values.values() if isinstance(values, dict) else values
        ranking = {v: i for i, v in enumerate(sorted(set(items)))}
        if isinstance(values, dict):
            return {k: ranking[v] for k, v in values.items()}
        return [ranking[v] for v in values]

    arg_col = compress(arg_col)
    n_classes = -1
    for _ in range(len(rows) + len(arg_col) + 2):
        row_col = compress([tuple(arg_col[a] for a in r) for r in rows])
        occ: dict[str, list[tuple[int, int]]] = {a: [] for a in arg_col}
        for ir, r in enumerate(rows):
            for pos, a in enumerate(r):
                occ[a].append((row_col[ir], pos))
        arg_col = compress({a: (arg_col[a], tuple(sorted(occ[a]))) for a in arg_col})
        now = len(set(row_col)) + len(set(arg_col.values()))
        if now == n_classes:
            break
        n_classes = now
    classes: dict[int, list[int]] = {}
    for ir, c in enumerate(row_col):
        classes.setdefault(c, []).append(ir)
    return [classes[c] for c in sorted(classes)]


def _name_matrix(rows: list[list[str]]) -> tuple[tuple[tuple[int, ...], ...], dict[str, int]]:
    ids: dict[str, int] = {}
    mat = []
    for r in rows:
        mat.append(tuple(ids.setdefault(a, len(ids)) for a in r))
    return tuple(mat), ids


def _canonical_rows(
    einsum: BatchedEinsum, order: tuple[int, ...]
) -> tuple[tuple[tuple[int, ...], ...], list[int], dict[str, int]]:
    groups = _refine_rows(einsum, order)
    all_rows = [[row[k].name for k in order] for row in einsum.args]
    total = 1
    for g in groups:
        for m in range(2, len(g) + 1):
            total *= m
            if total > _MAX_TIE_PERMS:
                break
    if total > _MAX_TIE_PERMS:
        # too many tied orders to enumerate: order each colour class greedily --
        # always take the row that reads smallest under the names given so far
        # (unnamed arrays provisionally numbered in order of appearance)
        ids: dict[str, int] = {}
        row_order = []
        for g in groups:
            remaining = list(g)
            while remaining:
                def provisional(ir: int) -> tuple[int, ...]:
                    local = dict(ids)
                    return tuple(local.setdefault(a, len(local)) for a in all_rows[ir])

                pick = min(remaining, key=provisional)
                remaining.remove(pick)
                row_order.append(pick)
                for a in all_rows[pick]:
                    ids.setdefault(a, len(ids))
        mat, ids = _name_matrix([all_rows[ir] for ir in row_order])
        return mat, row_order, ids
    best: tuple[Any, list[int], dict[str, int]] | None = None
    for combo in product(*[list(permutations(g)) for g in groups]):
        row_order = [ir for part in combo for ir in part]
        mat, ids = _name_matrix([all_rows[ir] for ir in row_order])
        if best is None or mat < best[0]:
            best = (mat, row_order, ids)
    assert best is not None
    return best


def _canonical_form(
    einsum: BatchedEinsum,
) -> tuple[BatchedEinsum, dict[str, str]]:
    best_sig: Any = None
    best: Any = None
    for order in _candidate_orders(einsum):
        # index naming by first appearance: output, then operands in order
        idx_name: dict[str, str] = {}
        for idx in einsum.out_idx_set:
            idx_name.setdefault(idx, chr(ord("a") + len(idx_name)))
        for k in order:
            for idx in einsum.in_idx_sets[k]:
                idx_name.setdefault(idx, chr(ord("a") + len(idx_name)))
        if len(idx_name) > 26:
            raise NotImplementedError("more than 26 indices")
        subs = (
            tuple(tuple(idx_name[i] for i in einsum.in_idx_sets[k]) for k in order),
            tuple(idx_name[i] for i in einsum.out_idx_set),
        )
        # size parameters: numbered by the first canonical index carrying them
        param_rank: dict[str, str] = {}
        extents = []
        for old, new in sorted(idx_name.items(), key=lambda kv: kv[1]):
            ext = einsum.index_to_dim_length[old]
            if isinstance(ext, SizeParam):
                pname = param_rank.setdefault(ext.name, new.upper())
                extents.append((new, 1, pname))
            else:
                extents.append((new, 0, int(ext)))
        mat, row_order, ids = _canonical_rows(einsum, order)
        dtypes = tuple(
            np.dtype(einsum.arg_to_dtype[name]).name
            for name, _ in sorted(ids.items(), key=lambda kv: kv[1])
        )
        sig = (subs, tuple(extents), mat, dtypes)
        if best_sig is None or sig < best_sig:
            best_sig = sig
            best = (order, idx_name, param_rank, row_order, ids)

    order, idx_name, param_rank, row_order, ids = best

    def new_shape(shape: tuple[Any, ...]) -> tuple[Any, ...]:
        return tuple(
            SizeParam(param_rank[d.name]) if isinstance(d, SizeParam) else d for d in shape
        )

    new_args = tuple(
        tuple(
            Array(
                f"arg_{ids[einsum.args[ir][k].name]}",
                new_shape(einsum.args[ir][k].shape),
                einsum.args[ir][k].dtype,
            )
            for k in order
        )
        for ir in row_order
    )
    canon = BatchedEinsum(
        tuple(idx_name[i] for i in einsum.out_idx_set),
        tuple(tuple(idx_name[i] for i in einsum.in_idx_sets[k]) for k in order),
        new_args,
    )
    subst: dict[str, str] = dict(idx_name)
    subst.update({old: f"arg_{k}" for old, k in ids.items()})
    subst.update(param_rank)
    old_out, new_out = einsum.output_names, canon.output_names
    for new_pos, ir in enumerate(row_order):
        subst[old_out[ir]] = new_out[new_pos]
    return canon, subst


def canonicalize_einsum(einsum: BatchedEinsum) -> BatchedEinsum:
    """Canonical representative of *einsum*'s isomorphism class."""
    return _canonical_form(einsum)[0]


def get_substitution_mapping_between_isomorphic_batched_einsums(
    batched_einsum_from: BatchedEinsum, batched_einsum_to: BatchedEinsum
) -> Map[str, str]:
    """Entity renaming (indices, arrays, size parameters, outputs) that turns
    *batched_einsum_from* into *batched_einsum_to*; ``ValueError`` if they are
    not isomorphic (reference ``canonicalization.py:1100-1125``)."""
    canon_from, map_from = _canonical_form(batched_einsum_from)
    canon_to, map_to = _canonical_form(batched_einsum_to)
    if canon_from != canon_to:
        raise ValueError("Einsums are not isomorphic.")
    inv_to = {v: k for k, v in map_to.items()}
    return Map({old: inv_to[new] for old, new in map_from.items()})
