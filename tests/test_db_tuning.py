"""Facts database + autotuner host logic on CPU: key formats, record/query/retrieve, DB-hit
short-circuit, seeding, illegal configurations.  ``measure.timeit`` (the only thing that needs a
GPU) is replaced by a deterministic fake clock here; tests/test_gpu_measure.py runs the real one."""

import json
import os
import sqlite3

import numpy as np
import pytest

import feinsum_b200 as f
from feinsum_b200 import measure, sql_utils, tuning
from feinsum_b200.cl_utils import FakeCLDevice
from feinsum_b200.tuning import BoolParameter, IntParameter, TupleParameter, einsum_arg, transform_param
from tests import einsums as E

IMPLS = tuning._get_impls_path()


class FakeQueue:
    def __init__(self, name="NVIDIA B200"):
        self.device = FakeCLDevice(name)


@pytest.fixture()
def fake_clock(monkeypatch):
    calls = []

    def fake_timeit(einsum, *, transform, cq, long_dim_length=100_000, schedule=None):
        from feinsum_b200.codegen.cuda import generate_cuda

        prog = transform(generate_cuda(einsum), insn_match=None, kernel_name=None)
        params = dict(prog.params)
        warps = params.get("threads", 256) // 32
        if warps > 12:
            raise f.InvalidParameterError("shared memory exceeded")
        calls.append(params)
        return 1e-3 * (1.0 + abs(warps - 10) * 0.05)      # best at 10 warps

    monkeypatch.setattr(measure, "timeit", fake_timeit)
    return calls


def test_parameter_classes():
    with pytest.raises(TypeError):
        IntParameter(1.5, 3)
    with pytest.raises(TypeError):
        IntParameter(1, "E")
    assert tuning._points(IntParameter(2, 4)) == [2, 3, 4]
    assert tuning._points(BoolParameter()) == [False, True]
    tp = tuning._convert_to_tuning_param((IntParameter(8, 9), (BoolParameter(), IntParameter(0, 1))))
    assert isinstance(tp, TupleParameter)
    assert len(tuning._points(tp)) == 2 * 2 * 2 and (8, (False, 1)) in tuning._points(tp)


def test_decorators_and_bind_args():
    @einsum_arg("ndof", lambda e: e.shape[-1])
    @transform_param("warps", lambda e: IntParameter(4, 6))
    def transform(program, ndof, warps, insn_match=None, kernel_name=None):
        return (program, ndof, warps)

    assert [p.var_name for p in transform.transform_params] == ["warps"]
    assert [a.var_name for a in transform.einsum_derivative_args] == ["ndof"]
    bound = transform.bind_args(E.div(), warps=5)
    assert bound("prog", insn_match=None, kernel_name=None) == ("prog", 35, 5)
    assert transform.parameter_space(E.div()) == {"warps": [4, 5, 6]}


def test_key_encodings_match_the_reference_format():
    c = f.canonicalize_einsum(E.grad())
    assert json.loads(sql_utils.dump_index_to_length(c)) == {k: v for k, v in zip("abcde", [3, None, 35, 3, 35]) if v} \
        or set(json.loads(sql_utils.dump_index_to_length(c)).values()) == {3, 35}
    assert "e" not in json.loads(sql_utils.dump_index_to_length(E.grad()))       # symbolic axis omitted
    assert json.loads(sql_utils.dump_arg_names(c)) == [["arg_0", "arg_1", "arg_2"]]
    assert set(json.loads(sql_utils.dump_arg_to_dtype(c)).values()) == {"float64"}
    assert sql_utils.dump_device_name(FakeCLDevice("NVIDIA B200")) == "NVIDIA_B200"
    assert sql_utils.dump_device_name(FakeCLDevice("Intel(R) Xeon(R) CPU E5-2650 v4 @ 2.20GHz")) == \
        "Intel_R__Xeon_R__CPU_E5_2650_v4_AT_2DOT20GHz"
    assert sql_utils.load_transform_params('{"wg": [8, [1, 2]], "n": 3}') == {"wg": (8, (1, 2)), "n": 3}
    assert json.loads(sql_utils.dump_op_info(E.grad(), 100_000)) == {"float64": pytest.approx(0.798)}


def test_record_query_retrieve(tmp_path, fake_clock):
    db = str(tmp_path / "facts.sqlite")
    cq = FakeQueue()
    mod = os.path.join(IMPLS, "xre_rij_ej_to_xei.py")
    with pytest.raises(RuntimeError):
        sqlite3.connect(db).close()
        f.query(E.grad(), cq.device, database=db)                  # no facts table yet
    for warps in (8, 10, 12):
        f.record_facts(E.grad(), cq, mod, {"warps": warps, "variant": 1}, database=db, long_dim_length=1000)
    # look-up is invariant to renaming / operand order
    transposed = f.einsum("abn,bqp,nq->anp", f.array("Jac", (3, 3, "N")), f.array("Dmat", (3, 35, 35)),
                          f.array("v", ("N", 35)))
    assert not f.query(transposed, cq.device, database=db)         # D accessed transposed -> different einsum
    same = f.einsum("abn,nq,bpq->anp", f.array("Jac", (3, 3, "N")), f.array("v", ("N", 35)),
                    f.array("Dmat", (3, 35, 35)))                  # renamed + operands re-ordered
    facts = f.query(same, cq.device, database=db)
    assert len(facts) == 3 and {q.transform_params["warps"] for q in facts} == {8, 10, 12}
    assert all(q.transform_id == "xre_rij_ej_to_xei.py" and q.n_elements == 1000 for q in facts)
    assert facts[0].giga_op_rate("float64") == pytest.approx(7980 * 1000 * 1e-9 / facts[0].runtime_in_sec)
    assert not f.query(E.grad(), FakeCLDevice("NVIDIA H200"), database=db)
    with pytest.raises(f.NoFactInDatabaseError):
        f.query(E.div(), cq.device, database=db, err_if_no_results=True)
    best = f.retrieve(same, cq.device, database=db)
    from feinsum_b200.codegen.cuda import generate_cuda
    assert dict(best(generate_cuda(E.grad()), insn_match=None, kernel_name=None).params) == {"threads": 320, "variant": 1}
    only8 = f.retrieve(same, cq.device, database=db, consider_query=lambda q: q.transform_params["warps"] == 8)
    assert dict(only8(generate_cuda(E.grad())).params)["threads"] == 256
    timed = f.get_timed_einsums_in_db(cq.device, database=db)
    assert timed == (f.canonicalize_einsum(E.grad()),)


def test_autotune_searches_records_and_short_circuits(tmp_path, fake_clock):
    db = str(tmp_path / "facts.sqlite")
    cq = FakeQueue()
    mod = os.path.join(IMPLS, "xre_rij_xej_to_ei.py")
    with pytest.raises(ValueError):
        f.autotune(E.div(), "relative/path.py", cq, db_path=db)
    best = f.autotune(E.div(), mod, cq, db_path=db, long_dim_length=1000)
    assert best == {"warps": 10, "variant": 1, "ctas_per_sm": 0, "tile_e8": 0}   # p = 4: grid / tile knobs are fixed
    assert len(fake_clock) == 5                                   # warps 8..12 timed, 13..16 illegal
    assert len(f.query(E.div(), cq.device, database=db)) == 5     # illegal points are not recorded
    fake_clock.clear()
    best2 = f.autotune(E.div(), mod, cq, db_path=db, long_dim_length=1000)
    assert best2 == best and len(fake_clock) == 0                 # every legal point was a DB hit
    # test_limit bounds the number of NEW trials
    db2 = str(tmp_path / "facts2.sqlite")
    f.autotune(E.div(), mod, cq, db_path=db2, long_dim_length=1000, test_limit=2)
    assert len(fake_clock) <= 2


def test_autotune_tuple_parameter_module(tmp_path, monkeypatch):
    """Reference test/tuning_impls_tests/test_tuple_args.py: a TupleParameter space end to end."""
    mod = tmp_path / "tuple_space.py"
    mod.write_text(
        "from feinsum_b200.tuning import IntParameter, transform_param\n"
        "@transform_param('wg_size', lambda ensm: (IntParameter(8, 9), IntParameter(8, 10)))\n"
        "def transform(program, wg_size, insn_match=None, kernel_name=None):\n"
        "    assert isinstance(wg_size, tuple) and len(wg_size) == 2\n"
        "    return program.with_params(tile_e=wg_size[0] * wg_size[1])\n")
    seen = []

    def fake_timeit(einsum, *, transform, cq, long_dim_length=100_000, schedule=None):
        from feinsum_b200.codegen.cuda import generate_cuda

        tile = dict(transform(generate_cuda(einsum)).params)["tile_e"]
        seen.append(tile)
        return 1e-3 / tile

    monkeypatch.setattr(measure, "timeit", fake_timeit)
    expr = f.einsum("ijk->ij", f.array("P", ("I", 72, 4), np.float64))
    db = str(tmp_path / "t.sqlite")
    best = f.autotune(expr, str(mod), FakeQueue(), db_path=db, long_dim_length=100)
    assert best == {"wg_size": (9, 10)} and len(seen) == 6
    facts = f.query(expr, FakeCLDevice("NVIDIA B200"), database=db)
    assert {q.transform_params["wg_size"] for q in facts} == {(a, b) for a in (8, 9) for b in (8, 9, 10)}
    assert all(os.path.isabs(q.transform_id) for q in facts)      # module outside tuning/impls -> absolute path


def test_shipped_database_has_b200_facts():
    """The package ships ``feinsum_b200/data/cuda_facts_v1.sqlite`` (filled by ``tools/populate_db.py`` on a
    B200: autotuner over the BASELINE einsums at E = 4 000 000 and 100 000) -- the counterpart of the
    reference's ``data/transform_archive_v5.sqlite``; ``retrieve`` with the default database must hand back
    a launch configuration for every BASELINE einsum, found under any renaming of it."""
    from feinsum_b200.codegen.cuda import generate_cuda

    dev = FakeCLDevice("NVIDIA B200")
    assert os.path.exists(f.DEFAULT_DB)
    for dt in ("float64", "float32"):
        for e in (E.grad(dtype=dt), E.div(dtype=dt), E.lift_fe(dtype=dt), E.lift_ef(dtype=dt), E.tensor_product(0, 8, dt)):
            facts = f.query(e, dev)
            assert facts and {q.n_elements for q in facts} == {100_000, 4_000_000}
            best = f.retrieve(e, dev)
            assert best(generate_cuda(e)).kernel_id == generate_cuda(e).kernel_id
    fp32_best = min((q for q in f.query(E.div(dtype="float32"), dev) if q.n_elements == 4_000_000),
                    key=lambda q: q.runtime_in_sec)
    assert fp32_best.transform_params["variant"] == 3          # tcgen05 wins the fp32 search
    renamed = f.einsum("abn,nq,bpq->anp", f.array("Jac", (3, 3, "N")), f.array("v", ("N", 35)), f.array("Dmat", (3, 35, 35)))
    assert len(f.query(renamed, dev)) == len(f.query(E.grad(), dev))
    assert len(f.get_timed_einsums_in_db(dev)) == 10


def test_launch_spaces_depend_on_the_shape():
    """the lower orders tune over the grid and the tile, p = 4 over the warp count (and, for the
    fp64 gradient, the formulation)."""
    from feinsum_b200.tuning import get_transform_func_from_module_path

    pt = get_transform_func_from_module_path(os.path.join(IMPLS, "xre_rij_ej_to_xei.py"))
    p4 = pt.parameter_space(f.canonicalize_einsum(E.grad()))
    p2 = pt.parameter_space(f.canonicalize_einsum(E.grad(ndof=10)))
    p4f = pt.parameter_space(E.grad(dtype="float32"))
    assert p4["warps"] == list(range(8, 17)) and p4["ctas_per_sm"] == [0] and p4["tile_e8"] == [0]
    assert p4["formulation"] == [1, 2] and p4f["formulation"] == [1] and p4f["variant"] == [1, 2, 3]
    assert p2["warps"] == [8] and p2["ctas_per_sm"] == [0, 1, 2, 3, 4] and p2["formulation"] == [1]
    prog = pt.bind_args(E.grad(), warps=12, formulation=2)(f.generate_cuda(E.grad()))
    assert dict(prog.params) == {"threads": 384, "variant": 1, "stages": 2}
    lo = get_transform_func_from_module_path(os.path.join(IMPLS, "ifj_fe_fej_to_ei.py"))
    e = E.lift_fe(nvol=10, nfd=6)
    assert lo.parameter_space(e)["ctas_per_sm"] == [0, 1, 2, 3, 4]
    assert dict(lo.bind_args(e, warps=8, ctas_per_sm=2)(f.generate_cuda(e)).params) == \
        {"threads": 256, "variant": 1, "ctas_per_sm": 2}
