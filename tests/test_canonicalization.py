"""Canonical form / isomorphism of batched einsums (the database key).  Cases follow the
reference's own tests (reference test/test_feinsum.py:34-311): DG einsums under renaming,
automorphic operand positions, a 500-row einsum, and a renaming fuzz."""

import numpy as np
import pytest

import feinsum_b200 as f
from feinsum_b200.canonicalization import (
    canonicalize_einsum,
    get_substitution_mapping_between_isomorphic_batched_einsums,
)
from tests import fuzzlib


def iso(a, b):
    return canonicalize_einsum(a) == canonicalize_einsum(b)


def _div_components(sub, jac, mat, fields, dtype="float64", ndim=3, ndofs=35):
    return f.batched_einsum(sub, [[f.array(j, ("E", ndim), dtype), f.array(mat, (ndim, ndofs, ndofs), dtype),
                                   f.array(u, ("E", ndofs), dtype)] for j, u in zip(jac, fields)])


def test_dg_einsums_under_renaming():
    e1 = _div_components("es, sij, ej -> ei", ["Jx", "Jy", "Jz"], "R", ["ux", "uy", "uz"])
    e2 = _div_components("td, dkl, tl -> tk", ["Jacx", "Jacy", "Jacz"], "ref_mat", ["x_dofs", "y_dofs", "z_dofs"])
    e3 = _div_components("td, dkl, tl -> tk", ["Jacx", "Jacy", "Jacz"], "ref_mat", ["u", "u", "u"])
    e4 = _div_components("es, sij, ej -> ei", ["Jx", "Jy", "Jz"], "R", ["ux", "uy", "uz"], "float32")
    assert iso(e1, e2)
    assert iso(canonicalize_einsum(e1), canonicalize_einsum(e2))        # idempotent
    assert canonicalize_einsum(canonicalize_einsum(e1)) == canonicalize_einsum(e1)
    assert not iso(e2, e3)          # sharing pattern of the fields differs
    assert not iso(e1, e4)          # dtype is part of the key


def test_automorphic_operand_positions():
    A = lambda n, s, d="float64": f.array(n, s, d)  # noqa: E731
    assert iso(f.einsum("ij,ik->i", A("A", ("I", 10)), A("B", ("I", 10), "float32")),
               f.einsum("ik,ij->i", A("C", ("J", 10), "float32"), A("D", ("J", 10))))
    assert not iso(
        f.einsum("ijk,ij,ik->i", A("A", ("I", 10, 10)), A("B", ("I", 10)), A("C", ("I", 10), "float32")),
        f.einsum("ijk,ij,ik->i", A("A", ("I", 10, 10)), A("B", ("I", 10), "float32"), A("C", ("I", 10))))
    assert iso(
        f.einsum("ijk,ij,ik->i", A("A", ("I", 10, 10)), A("B", ("I", 10)), A("C", ("I", 10))),
        f.einsum("ijk,ik,ij->i", A("P", ("J", 10, 10)), A("Q", ("J", 10)), A("R", ("J", 10))))
    four = lambda sub, n, p: f.batched_einsum(  # noqa: E731
        sub, [[A(n[0], (p, 10, 10)), A(n[1], (p, 10)), A(n[2], (p, 10)), A(n[3], (p, 10))]])
    assert not iso(four("ijk,ik,ij,ij->i", "ABCD", "I"), four("ijk,ik,ij,ik->i", "PQRS", "L"))
    assert iso(four("ijk,ik,ij,ij->i", "ABCD", "I"), four("ikj,ik,ij,ik->i", "PQRS", "L"))
    two_rows = lambda sub, p, rows: f.batched_einsum(  # noqa: E731
        sub, [[A(r[0], (p, 10, 10)), A(r[1], (p, 10)), A(r[2], (p, 10)), A(r[3], (p, 10))] for r in rows])
    assert iso(two_rows("ijk,ik,ij,ij->i", "I", ["ABCD", "ABCB"]),
               two_rows("elm,em,el,el->e", "J", ["PQRQ", "PQRS"]))


def test_large_batch():
    e1 = f.batched_einsum("ij,ej->ei", [[f.array(f"u{i}", (35, 35)), f.array(f"v{i}", ("E", 35))] for i in range(500)])
    e2 = f.batched_einsum("et,st->es", [[f.array(f"a{i}", ("E", 35)), f.array(f"b{i}", (35, 35))] for i in range(500)])
    assert iso(e1, e2)


def test_canonical_names_and_axis_order_significance():
    grad = f.einsum("xre,rij,ej->xei", f.array("J", (3, 3, "E")), f.array("D", (3, 35, 35)), f.array("u", ("E", 35)))
    c = canonicalize_einsum(grad)
    assert sorted(c.all_args) == ["arg_0", "arg_1", "arg_2"]
    assert set(c.all_indices) == set("abcde")
    # transposing an operand's axes changes the memory layout -> a different einsum
    grad_t = f.einsum("xre,rji,ej->xei", f.array("J", (3, 3, "E")), f.array("D", (3, 35, 35)), f.array("u", ("E", 35)))
    assert not iso(grad, grad_t)
    # lift in its two layouts is two different keys
    from tests import einsums as E
    assert not iso(E.lift_ef(), E.lift_fe())
    assert iso(E.lift_fe(), E.lift_fe())


def test_substitution_mapping():
    e1 = _div_components("es, sij, ej -> ei", ["Jx", "Jy", "Jz"], "R", ["ux", "uy", "uz"])
    e2 = _div_components("td, dkl, tl -> tk", ["Jacx", "Jacy", "Jacz"], "ref_mat", ["x_dofs", "y_dofs", "z_dofs"])
    m = get_substitution_mapping_between_isomorphic_batched_einsums(e1, e2)
    assert m["e"] == "t" and m["s"] == "d" and m["i"] == "k" and m["j"] == "l"
    assert m["R"] == "ref_mat" and m["E"] == "E"
    assert {m["Jx"], m["Jy"], m["Jz"]} == {"Jacx", "Jacy", "Jacz"}
    # fields follow their Jacobians row by row
    pair = {m[f"J{c}"][-1]: m[f"u{c}"][0] for c in "xyz"}
    assert all(k == v for k, v in pair.items())
    e3 = _div_components("td, dkl, tl -> tk", ["Jacx", "Jacy", "Jacz"], "ref_mat", ["u", "u", "u"])
    with pytest.raises(ValueError):
        get_substitution_mapping_between_isomorphic_batched_einsums(e1, e3)


def test_fuzz_renaming_invariance():
    rng = np.random.default_rng(0)
    for _ in range(256):
        e = fuzzlib.random_batched_einsum(rng)
        assert canonicalize_einsum(e) == canonicalize_einsum(fuzzlib.shuffled_copy(e, rng))
