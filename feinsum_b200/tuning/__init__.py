"""
Autotuner over CUDA launch configurations.

Keeps the reference's vocabulary (reference ``src/feinsum/tuning/__init__.py:53-633``):
a *transform space* is a Python module exporting ``transform(program, <params>,
insn_match=None, kernel_name=None)``, decorated with :class:`transform_param`
(tunable: :class:`IntParameter`, :class:`BoolParameter`, tuples thereof) and
:class:`einsum_arg` (value derived from the einsum); :func:`autotune` walks the
space, times every point with :func:`feinsum_b200.measure.timeit` and records it in
the facts database, skipping points already recorded ("DB hit") and seeding the
order with recorded ones.  A point outside a kernel's legal space comes back from
the C ABI as ``FNSM_E_BAD_CONFIG`` -> :class:`InvalidParameterError` -> ``inf``.

The reference drives OpenTuner's bandit search; OpenTuner is not available here
and the CUDA spaces are small (tens of points), so the search is: every recorded
point first, then the whole grid when it has at most ``max_grid`` points, else
random samples of it -- until ``test_limit`` trials or ``stop_after`` seconds.
"""

from __future__ import annotations

import abc
import itertools
import logging
import os
import time
from collections.abc import Callable, Iterator, Mapping
from dataclasses import dataclass
from functools import cache, partial
from typing import Any

import numpy as np

from feinsum_b200.diagnostics import InvalidParameterError, TransformValidationError
from feinsum_b200.einsum import INT_CLASSES, BatchedEinsum

logger = logging.getLogger(__name__)


# {{{ parameter space


class TuningParameter(abc.ABC):  # noqa: B024
    """Parameter space of a launch configuration (abstract)."""


@dataclass(frozen=True, init=False)
class IntParameter(TuningParameter):
    """Integers in ``[low, high]``."""

    low: int
    high: int

    def __init__(self, low: Any, high: Any):
        if not isinstance(low, INT_CLASSES):
            raise TypeError("low must be an integer")
        if not isinstance(high, INT_CLASSES):
            raise TypeError("high must be an integer")
        object.__setattr__(self, "low", int(low))
        object.__setattr__(self, "high", int(high))


@dataclass(frozen=True)
class BoolParameter(TuningParameter):
    """*True* or *False*."""


@dataclass(frozen=True)
class TupleParameter(TuningParameter):
    """Cartesian product of the element spaces."""

    _data: tuple[TuningParameter, ...]


ConvertibleToTuningParamT = Any


def _convert_to_tuning_param(param: Any) -> TuningParameter:
    if isinstance(param, TuningParameter):
        return param
    if isinstance(param, tuple):
        return TupleParameter(tuple(_convert_to_tuning_param(el) for el in param))
    raise TypeError("Only instances of ConvertibleToTuningParamT are supported.")


def _points(param: TuningParameter) -> list[Any]:
    if isinstance(param, IntParameter):
        return list(range(param.low, param.high + 1))
    if isinstance(param, BoolParameter):
        return [False, True]
    if isinstance(param, TupleParameter):
        return [tuple(c) for c in itertools.product(*[_points(p) for p in param._data])]
    raise NotImplementedError(type(param))


# }}}


# {{{ einsum_arg / transform_param


@dataclass(frozen=True, repr=True)
class einsum_arg:  # noqa: N801
    """Static argument of the transform, computed from the einsum."""

    var_name: str
    func: Callable[[BatchedEinsum], Any]

    def __call__(self, fn: Callable[..., Any]) -> "ParametrizedTransform":
        if isinstance(fn, ParametrizedTransform):
            return ParametrizedTransform(fn.transform, (self, *fn.einsum_derivative_args), fn.transform_params)
        return ParametrizedTransform(fn, (self,), ())


@dataclass(frozen=True, repr=True)
class transform_param:  # noqa: N801
    """Tunable argument of the transform and its space for a given einsum."""

    var_name: str
    func: Callable[[BatchedEinsum], Any]

    def __call__(self, fn: Callable[..., Any]) -> "ParametrizedTransform":
        if isinstance(fn, ParametrizedTransform):
            return ParametrizedTransform(fn.transform, fn.einsum_derivative_args, (self, *fn.transform_params))
        return ParametrizedTransform(fn, (), (self,))


@dataclass(frozen=True, repr=True)
class ParametrizedTransform:
    transform: Callable[..., Any]
    einsum_derivative_args: tuple[einsum_arg, ...]
    transform_params: tuple[transform_param, ...]

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        return self.transform(*args, **kwargs)

    def bind_args(self, einsum: BatchedEinsum, **transform_args: Any) -> Any:
        """Bind the einsum-derived and the given tunable arguments -> a ``TransformT``."""
        return partial(self.transform,
                       **{arg.var_name: arg.func(einsum) for arg in self.einsum_derivative_args},
                       **transform_args)

    def parameter_space(self, einsum: BatchedEinsum) -> dict[str, list[Any]]:
        return {p.var_name: _points(_convert_to_tuning_param(p.func(einsum))) for p in self.transform_params}


# }}}


@cache
def _get_impls_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "impls")


class ConfigurationNotInDBError(LookupError):
    pass


def get_transform_func_from_module_path(module_path: str) -> ParametrizedTransform:
    from importlib import util

    _, filename = os.path.split(module_path)
    assert filename.endswith(".py")
    spec = util.spec_from_file_location(filename[:-3], module_path)
    if spec is None or spec.loader is None:
        raise RuntimeError(f"Could not import 'transform' function from {module_path}.")
    module = util.module_from_spec(spec)
    spec.loader.exec_module(module)
    obj = module.transform
    if isinstance(obj, ParametrizedTransform):
        return obj
    assert callable(obj)
    return ParametrizedTransform(obj, (), ())


def _freeze(v: Any) -> Any:
    return tuple(_freeze(k) for k in v) if isinstance(v, (list, tuple)) else v


def _iter_configs(space: dict[str, list[Any]], seeds: list[dict[str, Any]], rng: np.random.Generator,
                  max_grid: int) -> Iterator[dict[str, Any]]:
    names = sorted(space)
    seen: set[tuple[Any, ...]] = set()

    def key(cfg: Mapping[str, Any]) -> tuple[Any, ...]:
        return tuple(_freeze(cfg[n]) for n in names)

    for cfg in seeds:                                 # recorded points first (reference :418-469)
        if set(cfg) == set(names) and all(_freeze(cfg[n]) in space[n] for n in names) and key(cfg) not in seen:
            seen.add(key(cfg))
            yield dict(cfg)
    total = int(np.prod([len(space[n]) for n in names], dtype=np.float64)) if names else 1
    if total <= max_grid:
        grid = [dict(zip(names, combo)) for combo in itertools.product(*[space[n] for n in names])]
        rng.shuffle(grid)                             # type: ignore[arg-type]
        for cfg in grid:
            if key(cfg) not in seen:
                seen.add(key(cfg))
                yield cfg
    else:
        misses = 0
        while len(seen) < total and misses < 1000:
            cfg = {n: space[n][int(rng.integers(0, len(space[n])))] for n in names}
            if key(cfg) in seen:
                misses += 1
                continue
            seen.add(key(cfg))
            yield cfg


def autotune(einsum: BatchedEinsum, module_path: str, cq: Any, *, db_path: str | None = None,
             long_dim_length: int = 100_000, test_limit: int | None = None, stop_after: float | None = None,
             skip_value_mismatch: bool = False, max_grid: int = 4096, seed: int = 0) -> dict[str, Any] | None:
    """Search the launch-configuration space of *module_path* for *einsum* on *cq*, recording every
    timed point in *db_path* (same signature as the reference's ``autotune``; returns the best
    configuration found, which the reference leaves to a later ``retrieve``)."""
    from feinsum_b200 import sql_utils
    from feinsum_b200.canonicalization import canonicalize_einsum

    if not os.path.isabs(module_path):
        raise ValueError("autotune expects an absolute path for the module")
    if db_path is None:
        db_path = sql_utils.DEFAULT_DB
    einsum = canonicalize_einsum(einsum)
    ptransform = get_transform_func_from_module_path(module_path)
    space = ptransform.parameter_space(einsum)

    dirpath, transform_id = os.path.split(module_path)
    if os.path.abspath(dirpath) != _get_impls_path():
        transform_id = module_path
    recorded: dict[tuple[Any, ...], float] = {}
    names = sorted(space)
    try:
        for q in sql_utils.query(einsum, cq.device, database=db_path):
            if q.transform_id == transform_id and set(q.transform_params) == set(names) \
                    and q.n_elements in (0, long_dim_length):
                recorded[tuple(_freeze(q.transform_params[n]) for n in names)] = q.runtime_in_sec
    except (RuntimeError, OSError):
        pass                                          # no table / no file yet
    seeds = [dict(zip(names, k)) for k, _ in sorted(recorded.items(), key=lambda kv: kv[1])]

    t_start = time.time()
    n_trials = 0
    best: tuple[float, dict[str, Any]] | None = None
    for cfg in _iter_configs(space, seeds, np.random.default_rng(seed), max_grid):
        if test_limit is not None and n_trials >= test_limit:
            break
        if stop_after is not None and time.time() - t_start > stop_after:
            break
        k = tuple(_freeze(cfg[n]) for n in names)
        if k in recorded:
            logger.info("DB Hit for %s", cfg)
            runtime = recorded[k]
        else:
            n_trials += 1
            try:
                runtime = sql_utils.record_facts(einsum, cq, module_path, cfg, db_path, long_dim_length)
            except InvalidParameterError as err:
                logger.info("Ignored configuration %s due to %s", cfg, err)
                runtime = float("inf")
            except TransformValidationError:
                if not skip_value_mismatch:
                    raise
                logger.info("Ignored configuration %s due to a value mismatch", cfg)
                runtime = float("inf")
        if best is None or runtime < best[0]:
            best = (runtime, cfg)
    return None if best is None or not np.isfinite(best[0]) else best[1]
