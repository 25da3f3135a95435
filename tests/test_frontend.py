"""Front-end parity with the reference: every expectation here was computed by
the reference's own ``einsum.py`` / ``make_einsum.py`` (tests/golden/make_golden.py)."""

import json
import os

import numpy as np
import pytest

import feinsum_b200 as f
from feinsum_b200.einsum import SizeParam
from tests.einsums import from_spec

EXC = {"ValueError": ValueError, "TypeError": TypeError, "NotImplementedError": NotImplementedError}


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "frontend.json")) as fh:
        return json.load(fh)


def _dec(d):
    return SizeParam(d["param"]) if isinstance(d, dict) else d


def test_valid_constructions_match_reference(golden):
    assert len(golden["valid"]) >= 10
    for name, exp in golden["valid"].items():
        e = from_spec(exp["spec"])
        assert e.get_subscripts() == exp["get_subscripts"], name
        assert (e.b, e.n, e.ndim) == (exp["b"], exp["n"], exp["ndim"]), name
        assert e.shape == tuple(_dec(d) for d in exp["shape"]), name
        assert dict(e.index_to_dim_length) == {
            k: _dec(v) for k, v in exp["index_to_dim_length"].items()
        }, name
        assert {k: tuple(v) for k, v in e.arg_to_shape.items()} == {
            k: tuple(_dec(d) for d in v) for k, v in exp["arg_to_shape"].items()
        }, name
        assert {k: np.dtype(v).name for k, v in e.arg_to_dtype.items()} == exp["arg_to_dtype"]
        assert list(e.sum_indices) == exp["sum_indices"], name
        assert sorted(e.all_args) == exp["all_args"]
        assert sorted(e.all_indices) == exp["all_indices"]
        assert sorted(p.name for p in e.all_size_params) == exp["all_size_params"]
        got_access = {
            k: [type(v).__name__, getattr(v, "output_index", getattr(v, "index", -1))]
            for k, v in e.index_to_access_descr.items()
        }
        assert got_access == exp["access"], name


def test_invalid_constructions_raise_like_reference(golden):
    for name, exp in golden["invalid"].items():
        assert exp["raises"] is not None, name
        with pytest.raises(EXC[exp["raises"]]):
            from_spec(exp["spec"])


def test_ellipsis_satisfies_documented_and_actual_behaviour():
    # reference make_einsum.py:98 documents NotImplementedError, raises TypeError
    with pytest.raises(NotImplementedError):
        f.einsum("...j,j->...", f.array("A", (3, 4)), f.array("x", 4))


def test_bad_shape_components(golden):
    vals = {"negative": -1, "float": 2.5, "inf": np.inf, "none": None}
    for name, exc in golden["bad_shape_component"].items():
        with pytest.raises(EXC[exc]):
            f.array("A", (3, vals[name]))


def test_array_basics():
    a = f.array("A", 4, "float32")
    assert a.shape == (4,) and a.ndim == 1 and a.dtype == np.float32
    b = f.array("B", ("E", np.int64(3)))
    assert b.shape == (SizeParam("E"), 3) and b.dtype == np.float64
    c = b.copy(name="C")
    assert c.name == "C" and c.shape == b.shape and b.name == "B"
    with pytest.raises(Exception):
        b.name = "x"  # frozen


def test_batched_einsum_is_hashable_and_comparable():
    from tests.einsums import grad

    assert grad() == grad() and hash(grad()) == hash(grad())
    assert grad() != grad("float32")
    e = grad()
    assert e.copy() == e
    assert e.copy(out_idx_set=("x", "e", "i")) == e


def test_output_names_and_str():
    from tests.einsums import lift_ef

    e = lift_ef()
    assert e.output_names == ("_fe_out", "_fe_out_0", "_fe_out_1", "_fe_out_2")
    s = str(e)
    assert "DOMAINS" in s and "_fe_out_2[e, i]" in s and "0 <= e < E" in s
    assert "J: float64" in s


def test_abstract_axis_access():
    with pytest.raises(TypeError):
        f.EinsumAxisAccess()
    assert f.FreeAxis(1) == f.FreeAxis(1) and f.SummationAxis(0) != f.SummationAxis(1)


def test_index_name_generator():
    g = f.IndexNameGenerator(frozenset({"c"}))
    assert [g(), g(), g()] == ["a", "b", "d"]
    g2 = f.IndexNameGenerator()
    for _ in range(26):
        g2()
    with pytest.raises(RuntimeError):
        g2()


def test_size_param_division_is_undefined():
    with pytest.raises(TypeError):
        SizeParam("E") / 4


def test_immutable_map():
    from feinsum_b200._immutable import Map

    m = Map(a=1, b=2)
    assert m["a"] == 1 and len(m) == 2 and hash(m) == hash(Map(b=2, a=1))
    m2 = m.update({"c": 3})
    assert "c" not in m and m2["c"] == 3
    assert m.set("a", 5)["a"] == 5 and "a" not in m.delete("a")
    with pytest.raises(AttributeError):
        m.x = 1


def test_tccg_benchmark_getter():
    """reference test/test_feinsum.py:286-288, plus every field against what the reference's own getter
    builds (tests/golden/tccg.json, written by make_golden.py from src/feinsum/utils.py:103-233)."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "tccg.json")) as fh:
        cases = json.load(fh)
    assert [c["i"] for c in cases] == list(range(1, 49))
    for c in cases:
        e = f.utils.get_tccg_benchmark(c["i"])
        assert isinstance(e, f.BatchedEinsum)
        assert e.get_subscripts() == c["subscripts"]
        assert [a.name for a in e.args[0]] == c["arg_names"]
        assert [[int(d) for d in e.arg_to_shape[n]] for n in c["arg_names"]] == c["arg_shapes"]
        assert [int(d) for d in e.shape] == c["shape"]
        assert all(np.dtype(dt) == np.float64 for dt in e.arg_to_dtype.values())
    assert np.dtype(f.utils.get_tccg_benchmark(3, "float32").arg_to_dtype["A"]) == np.float32
    for bad in (0, 49, -1):
        with pytest.raises(ValueError):
            f.utils.get_tccg_benchmark(bad)
