"""
Contraction schedules: in which order the operands of an einsum are multiplied.

Mirrors the reference's ``feinsum.contraction_schedule`` (reference
``src/feinsum/contraction_schedule.py:27-178``): the same
:class:`ContractionSchedule` record and the two constructors
:func:`get_trivial_contraction_schedule` and
:func:`get_opt_einsum_contraction_schedule`.

The reference delegates the second one to ``opt_einsum.contract_path(...,
optimize="optimal")``, which is not installed here.  Its *published*
algorithm is restated instead: exhaustive depth-first search over pairwise
contractions minimising the summed flop estimate ``prod(extents of all
indices touched) * (1 + [an index is summed away])``, symbolic extents
replaced by ``long_dim_length`` (10**6 by default, reference
``contraction_schedule.py:135``).  Spelling conventions that the rest of the
code base (and the reference's tests / database) observe are kept:

* a step lists the contracted operands from the highest current position to
  the lowest (``"ej,rij->rie"`` for the DG gradient),
* an intermediate's indices are ordered by ``(extent, letter)``,
* intermediates are named ``_fe_tmp``, ``_fe_tmp_0``, ...; the last result
  is ``_fe_out``.

The schedule fixes the FLOP count every GFLOP/s figure is quoted on
(reference ``measure.py:278-331``); see :func:`feinsum_b200.measure.get_flops_per_dtype`.
"""

from __future__ import annotations

from dataclasses import dataclass, replace
from itertools import combinations
from typing import Any

from feinsum_b200.einsum import BatchedEinsum, SizeParam


class Argument:
    """Abstract operand of one schedule step."""

    def __init__(self) -> None:
        if type(self) is Argument:
            raise TypeError(
                "Argument is abstract and cannot be instantiated directly."
            )


@dataclass(frozen=True)
class IntermediateResult(Argument):
    """Result of an earlier step, referred to by name."""

    name: str


@dataclass(frozen=True, eq=True, repr=True)
class EinsumOperand(Argument):
    """The ``ioperand``-th operand of the parent einsum."""

    ioperand: int


@dataclass(frozen=True, eq=True, repr=True)
class ContractionSchedule:
    """Step ``i`` evaluates ``subscripts[i]`` on ``arguments[i]`` and names the
    result ``result_names[i]``."""

    subscripts: tuple[str, ...]
    result_names: tuple[str, ...]
    arguments: tuple[tuple[Argument, ...], ...]

    def __post_init__(self) -> None:
        assert len(self.subscripts) == len(self.result_names) == len(self.arguments)

    @property
    def nsteps(self) -> int:
        return len(self.subscripts)

    def copy(self, **kwargs: Any) -> "ContractionSchedule":
        return replace(self, **kwargs)


def get_trivial_contraction_schedule(einsum: BatchedEinsum) -> ContractionSchedule:
    """Everything in one step (reference ``contraction_schedule.py:101-110``)."""
    return ContractionSchedule(
        (einsum.get_subscripts(),),
        ("_fe_out",),
        (tuple(EinsumOperand(i) for i in range(einsum.n)),),
    )


# {{{ optimal pairwise path


def _pair_cost(
    a: frozenset[str],
    b: frozenset[str],
    keep: frozenset[str],
    size: dict[str, int],
) -> tuple[int, frozenset[str]]:
    """Flop estimate and surviving indices for contracting operands a, b."""
    touched = a | b
    result = touched & keep
    cost = 1
    for idx in touched:
        cost *= size[idx]
    if touched - result:
        cost *= 2  # one multiply + one reduction add per iteration
    return cost, result


def _optimal_path(
    inputs: list[frozenset[str]], output: frozenset[str], size: dict[str, int]
) -> list[tuple[int, ...]]:
    """Exhaustive DFS over pairwise contraction orders, first-found minimum."""
    n = len(inputs)
    if n <= 2:
        return [tuple(range(n))]

    best: dict[str, Any] = {"cost": None, "path": None}

    def visit(
        remaining: list[frozenset[str]], path: list[tuple[int, ...]], cost: int
    ) -> None:
        if len(remaining) == 1:
            if best["cost"] is None or cost < best["cost"]:
                best["cost"], best["path"] = cost, list(path)
            return
        for i, j in combinations(range(len(remaining)), 2):
            others = [s for k, s in enumerate(remaining) if k not in (i, j)]
            keep = output.union(*others) if others else output
            step_cost, result = _pair_cost(remaining[i], remaining[j], keep, size)
            new_cost = cost + step_cost
            if best["cost"] is not None and new_cost >= best["cost"]:
                continue
            path.append((i, j))
            visit([*others, result], path, new_cost)
            path.pop()

    visit(list(inputs), [], 0)
    assert best["path"] is not None
    return best["path"]  # type: ignore[no-any-return]


# }}}


def _fresh_name(base: str, taken: set[str]) -> str:
    name, k = base, 0
    while name in taken:
        name = f"{base}_{k}"
        k += 1
    taken.add(name)
    return name


def get_opt_einsum_contraction_schedule(
    expr: BatchedEinsum, **opt_einsum_kwargs: Any
) -> ContractionSchedule:
    """
    Flop-optimal pairwise schedule (what ``opt_einsum`` calls
    ``optimize="optimal"``; reference ``contraction_schedule.py:113-178``).

    Accepted keyword: ``long_dim_length`` (value substituted for symbolic
    extents while costing, default ``1_000_000``).  ``optimize="optimal"`` and
    ``use_blas=False`` are accepted and are the only supported settings.
    """
    long_dim_length = int(opt_einsum_kwargs.pop("long_dim_length", 1_000_000))
    optimize = opt_einsum_kwargs.pop("optimize", "optimal")
    opt_einsum_kwargs.pop("use_blas", None)
    if optimize != "optimal":
        raise NotImplementedError(
            f"optimize={optimize!r}: only the exhaustive 'optimal' search is built in."
        )
    if opt_einsum_kwargs:
        raise TypeError(f"unexpected arguments: {sorted(opt_einsum_kwargs)}")

    size = {
        idx: long_dim_length if isinstance(ext, SizeParam) else int(ext)
        for idx, ext in expr.index_to_dim_length.items()
    }
    # operands as *ordered* index strings; sets are used only for costing
    operands: list[tuple[str, ...]] = [tuple(s) for s in expr.in_idx_sets]
    output = tuple(expr.out_idx_set)

    path = _optimal_path(
        [frozenset(s) for s in operands], frozenset(output), size
    )

    current_args: list[Argument] = [EinsumOperand(i) for i in range(expr.n)]
    current_idx: list[tuple[str, ...]] = list(operands)
    taken: set[str] = set()
    subscripts: list[str] = []
    result_names: list[str] = []
    arguments: list[tuple[Argument, ...]] = []

    for istep, positions in enumerate(path):
        picked = sorted(positions, reverse=True)
        is_last = istep == len(path) - 1
        in_strs = [current_idx[p] for p in picked]
        rest_idx = [s for k, s in enumerate(current_idx) if k not in positions]
        rest_args = [a for k, a in enumerate(current_args) if k not in positions]
        if is_last:
            result_idx = output
        else:
            touched = set().union(*in_strs)
            needed = set(output).union(*rest_idx)
            result_idx = tuple(
                sorted(touched & needed, key=lambda idx: (size[idx], idx))
            )
        subscripts.append(
            ",".join("".join(s) for s in in_strs) + "->" + "".join(result_idx)
        )
        arguments.append(tuple(current_args[p] for p in picked))
        result_names.append(_fresh_name("_fe_tmp", taken))
        current_idx = [*rest_idx, result_idx]
        current_args = [*rest_args, IntermediateResult(result_names[-1])]

    assert len(current_args) == 1
    result_names[-1] = _fresh_name("_fe_out", taken)
    return ContractionSchedule(
        tuple(subscripts), tuple(result_names), tuple(arguments)
    )
