"""
Database of timing facts: which CUDA launch configuration ran which einsum on
which device, how fast.

Same entry points as the reference's ``feinsum.sql_utils`` (reference
``src/feinsum/sql_utils.py:48-530``): :func:`query`, :func:`retrieve`,
:func:`record_facts`, :func:`get_timed_einsums_in_db`, ``DEFAULT_DB`` -- and the
same *key*: the canonicalised einsum (subscripts, fixed extents, operand-name
matrix, dtypes; reference ``sql_utils.py:56-132``) plus the sanitised device
name.  What a row stores differs: instead of a loopy transform script and its
parameters, ``transform_id`` names a launch-configuration module under
``feinsum_b200/tuning/impls`` and ``transform_params`` its CUDA tile / warp /
variant choices; the new columns keep what the B200 report needs (elements
timed, GFLOP/s, GB/s, fraction of the roofline).  Table
``FEINSUM_CUDA_FACTS``; the reference's ``FEINSUM_TIMING_FACTS`` rows (Titan V,
loopy) are not comparable and are not imported.  The shipped default database
(``data/cuda_facts_v1.sqlite``) holds the autotuner's B200 facts for the BASELINE
einsums (``tools/populate_db.py``).
"""

from __future__ import annotations

import json
import logging
import os
import sqlite3
from collections.abc import Callable, Mapping, Sequence
from dataclasses import dataclass
from functools import cached_property
from typing import Any

import numpy as np

from feinsum_b200._immutable import Map
from feinsum_b200.diagnostics import NoFactInDatabaseError
from feinsum_b200.einsum import INT_CLASSES, BatchedEinsum, SizeParam

logger = logging.getLogger(__name__)

DEFAULT_DB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "cuda_facts_v1.sqlite")
TIMINGS_TABLENAME = "FEINSUM_CUDA_FACTS"


# {{{ key encodings (identical to the reference's, sql_utils.py:56-132)


def dump_arg_to_dtype(einsum: BatchedEinsum) -> str:
    return json.dumps({arg: np.dtype(dtype).name for arg, dtype in einsum.arg_to_dtype.items()}, sort_keys=True)


def dump_index_to_length(einsum: BatchedEinsum) -> str:
    return json.dumps({k: int(v) for k, v in einsum.index_to_dim_length.items() if isinstance(v, INT_CLASSES)},
                      sort_keys=True)


def dump_arg_names(einsum: BatchedEinsum) -> str:
    return json.dumps([[arg.name for arg in row] for row in einsum.args])


def dump_device_name(device: Any) -> str:
    name = device.name
    assert isinstance(name, str)
    return (name.replace(" ", "_").replace("-", "_").replace("@", "AT").replace("(", "_")
            .replace(")", "_").replace(".", "DOT"))


def dump_compiler_version(device: Any) -> str:
    return f"{getattr(device, 'vendor', 'NVIDIA')}-{getattr(device, 'driver_version', 'unknown')}-sm_100a"


def dump_op_info(einsum: BatchedEinsum, long_dim_length: int) -> str:
    from feinsum_b200.measure import _get_giga_ops_from_einsum

    return json.dumps({np.dtype(k).name: v for k, v in _get_giga_ops_from_einsum(einsum, long_dim_length).items()},
                      sort_keys=True)


def load_op_info(op_info: str) -> Map[np.dtype[Any], float]:
    return Map({np.dtype(k): float(v) for k, v in json.loads(op_info).items()})


def _to_tuples(x: Any) -> Any:
    return tuple(_to_tuples(k) for k in x) if isinstance(x, list) else x


def load_transform_params(params_str: str) -> Map[str, Any]:
    return Map({k: _to_tuples(v) for k, v in json.loads(params_str).items()})


# }}}


@dataclass(frozen=True)
class QueryInfo:
    transform_id: str
    transform_params: Map[str, Any]
    runtime_in_sec: float
    compiler_version: str
    giga_op_info: Map[np.dtype[Any], float]
    _einsum: BatchedEinsum
    n_elements: int = 0
    roofline_frac: float | None = None

    def giga_op_rate(self, dtype: Any) -> float:
        return self.giga_op_info[np.dtype(dtype)] / self.runtime_in_sec

    @cached_property
    def transform(self) -> Any:
        from feinsum_b200.tuning import _get_impls_path, get_transform_func_from_module_path

        path = self.transform_id
        if not os.path.isabs(path):
            path = os.path.join(_get_impls_path(), path)
        return get_transform_func_from_module_path(path).bind_args(self._einsum, **self.transform_params)


def _connect(database: str | sqlite3.Connection) -> tuple[sqlite3.Connection, bool]:
    if isinstance(database, sqlite3.Connection):
        return database, False
    return sqlite3.connect(database), True


def _has_table(conn: sqlite3.Connection) -> bool:
    cur = conn.cursor()
    cur.execute("SELECT name FROM sqlite_master WHERE (type='table' AND name=?);", (TIMINGS_TABLENAME,))
    return bool(cur.fetchall())


def _create_timings_table_if_non_existent(conn: sqlite3.Connection) -> None:
    if not _has_table(conn):
        logger.info("Table %s not in DB, creating one.", TIMINGS_TABLENAME)
        conn.cursor().execute(
            f"CREATE TABLE {TIMINGS_TABLENAME} ("
            " ID INTEGER PRIMARY KEY AUTOINCREMENT,"
            " subscripts TEXT, index_to_length TEXT, args TEXT, arg_to_dtype TEXT, device_name TEXT,"
            " transform_id TEXT, transform_params TEXT, runtime_in_sec REAL, compiler_version TEXT,"
            " giga_op_info TEXT, timestamp TEXT,"
            " n_elements INTEGER, n_gpus INTEGER, gbytes REAL, achieved_gflops REAL, achieved_gbs REAL,"
            " roofline_frac REAL)")
    conn.commit()


def _key(einsum: BatchedEinsum, device: Any) -> tuple[str, str, str, str, str]:
    return (einsum.get_subscripts(), dump_index_to_length(einsum), dump_arg_names(einsum),
            dump_arg_to_dtype(einsum), dump_device_name(device))


def query(einsum: BatchedEinsum, cl_device: Any, *, database: str | sqlite3.Connection = DEFAULT_DB,
          err_if_no_results: bool = False) -> tuple[QueryInfo, ...]:
    """Facts of previously recorded runs of *einsum* on *cl_device* (anything with ``.name``)."""
    from feinsum_b200.canonicalization import canonicalize_einsum

    einsum = canonicalize_einsum(einsum)
    conn, own = _connect(database)
    try:
        if not _has_table(conn):
            raise RuntimeError(f"Database '{database}' does not contain the timing facts table.")
        cur = conn.cursor()
        cur.execute(
            "SELECT transform_id, transform_params, runtime_in_sec, compiler_version, giga_op_info,"
            f" n_elements, roofline_frac FROM {TIMINGS_TABLENAME} WHERE (subscripts = ? AND index_to_length = ?"
            " AND args = ? AND arg_to_dtype = ? AND device_name = ?);", _key(einsum, cl_device))
        facts = cur.fetchall()
    finally:
        if own:
            conn.close()
    result = tuple(
        QueryInfo(transform_id=f[0], transform_params=load_transform_params(f[1]), runtime_in_sec=f[2],
                  compiler_version=f[3], giga_op_info=load_op_info(f[4]), _einsum=einsum,
                  n_elements=int(f[5] or 0), roofline_frac=f[6])
        for f in facts)
    if not result and err_if_no_results:
        sizes = ", ".join(f"{idx}: {n}" for idx, n in einsum.index_to_dim_length.items()
                          if not isinstance(n, SizeParam))
        raise NoFactInDatabaseError(
            f"No facts found for the einsum: `{einsum.get_subscripts()} [{sizes}] [#outputs={einsum.b}]`.")
    return result


def retrieve(einsum: BatchedEinsum, cl_device: Any, *, database: str | sqlite3.Connection = DEFAULT_DB,
             consider_query: Callable[[QueryInfo], bool] | None = None) -> Any:
    """The launch configuration (as a bound transform) with the highest recorded
    op-throughput, summed over dtypes (reference ``sql_utils.py:247-294``)."""
    if consider_query is None:
        consider_query = lambda q: True  # noqa: E731
    queries = [q for q in query(einsum, cl_device, database=database, err_if_no_results=True) if consider_query(q)]
    if not queries:
        raise NoFactInDatabaseError(
            f"No facts found for the einsum: `{einsum}`, with the filtering function: {consider_query!r}.")
    best = max(queries, key=lambda q: sum(q.giga_op_rate(dt) for dt in q.giga_op_info))
    return best.transform


def _get_batched_einsum_from_sql_row(subscripts: str, index_to_length: Mapping[str, Any],
                                     arg_names: Sequence[Sequence[str]], arg_to_dtype: Mapping[str, str]) -> BatchedEinsum:
    from feinsum_b200.make_einsum import array, batched_einsum, parse_subscripts

    _, in_idx_sets = parse_subscripts(subscripts)
    lengths: dict[str, Any] = dict(index_to_length)
    for idx in {i for s in in_idx_sets for i in s}:
        if idx not in lengths:
            lengths[idx] = SizeParam(idx.upper())
    arg_to_shape = {arg: tuple(lengths[i] for i in idx_set)
                    for row in arg_names for idx_set, arg in zip(in_idx_sets, row)}
    rows = [[array(arg, arg_to_shape[arg], arg_to_dtype[arg]) for arg in row] for row in arg_names]
    return batched_einsum(subscripts, rows)


def get_timed_einsums_in_db(cl_device: Any, database: str | sqlite3.Connection = DEFAULT_DB) -> tuple[BatchedEinsum, ...]:
    """Every einsum with at least one fact on *cl_device*."""
    conn, own = _connect(database)
    try:
        cur = conn.cursor()
        cur.execute(f"SELECT subscripts, index_to_length, args, arg_to_dtype FROM {TIMINGS_TABLENAME}"
                    " WHERE device_name = ?;", (dump_device_name(cl_device),))
        facts = set(cur.fetchall())
    finally:
        if own:
            conn.close()
    seen = [_get_batched_einsum_from_sql_row(s, json.loads(i2l), json.loads(names), json.loads(a2d))
            for s, i2l, names, a2d in sorted(facts)]
    assert len(set(seen)) == len(seen)      # canonicalisation was sound
    return tuple(seen)


def _timestamp() -> str:
    from datetime import datetime, timezone

    try:
        from zoneinfo import ZoneInfo

        now = datetime.now(ZoneInfo("America/Chicago"))       # the reference stamps in Chicago time
    except Exception:  # noqa: BLE001
        now = datetime.now(timezone.utc)
    return now.strftime("%Y_%m_%d_%H%M%S")


def record_facts(einsum: BatchedEinsum, cq: Any, module_path: str, transform_params: Mapping[str, Any],
                 database: str | sqlite3.Connection = DEFAULT_DB, long_dim_length: int = 100_000) -> float:
    """Time *einsum* on *cq* with the launch configuration ``module_path(**transform_params)`` and
    store the fact (reference ``sql_utils.py:418-509``).  Returns the measured runtime in seconds."""
    from feinsum_b200 import measure
    from feinsum_b200.canonicalization import canonicalize_einsum
    from feinsum_b200.tuning import _get_impls_path, get_transform_func_from_module_path

    dirpath, transform_space_id = os.path.split(module_path)
    if os.path.abspath(dirpath) != _get_impls_path():
        transform_space_id = module_path
    einsum = canonicalize_einsum(einsum)
    transform = get_transform_func_from_module_path(module_path).bind_args(einsum, **transform_params)
    runtime = measure.timeit(einsum, cq=cq, transform=transform, long_dim_length=long_dim_length)
    device = cq.device
    logger.info("\n%s", measure._stringify_runtime_comparison_vs_roofline(
        einsum, runtime, device.name, long_dim_length=long_dim_length))

    giga_ops = measure._get_giga_ops_from_einsum(einsum, long_dim_length)
    gbytes = measure._get_footprint_gbytes(einsum, long_dim_length)
    try:
        roof = measure.get_roofline_flop_rate(einsum, device.name, long_dim_length)
        t_roof = max(giga_ops[dt] / roof[dt] for dt in giga_ops)
        frac = t_roof / runtime
    except Exception:  # noqa: BLE001  (unknown device: no peaks)
        frac = None
    conn, own = _connect(database)
    try:
        _create_timings_table_if_non_existent(conn)
        conn.cursor().execute(
            f"INSERT INTO {TIMINGS_TABLENAME} (subscripts, index_to_length, args, arg_to_dtype, device_name,"
            " transform_id, transform_params, runtime_in_sec, compiler_version, giga_op_info, timestamp,"
            " n_elements, n_gpus, gbytes, achieved_gflops, achieved_gbs, roofline_frac)"
            " VALUES (?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?)",
            (*_key(einsum, device), transform_space_id, json.dumps(dict(transform_params), sort_keys=True),
             runtime, dump_compiler_version(device), dump_op_info(einsum, long_dim_length), _timestamp(),
             int(long_dim_length), 1, gbytes, sum(giga_ops.values()) / runtime, gbytes / runtime, frac))
        conn.commit()
    finally:
        if own:
            conn.close()
    return runtime


def record_into_db(einsum: BatchedEinsum, cq: Any, module_path: str, transform_params: Mapping[str, Any],
                   database: str | sqlite3.Connection = DEFAULT_DB, long_dim_length: int = 100_000) -> float:
    import warnings

    warnings.warn("'record_into_db' is deprecated. Use 'record_facts' instead.", DeprecationWarning, stacklevel=2)
    return record_facts(einsum, cq, module_path, transform_params, database, long_dim_length)
