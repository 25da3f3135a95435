"""Einsum -> kernel family classification (renaming / operand-order invariant)."""

import numpy as np

import feinsum_b200 as f
from feinsum_b200.codegen import classify, generate_cuda, match_subscripts
from tests import einsums as E


def test_named_families():
    assert classify(E.grad()).kernel_id == "grad"
    assert classify(E.grad_batched(3)).kernel_id == "grad"
    assert classify(E.div()).kernel_id == "div"
    assert classify(E.lift_ef()).kernel_id == "lift_ef"
    assert classify(E.lift_fe()).kernel_id == "lift_fe"
    for mode in range(3):
        p = classify(E.tensor_product(mode))
        assert p.kernel_id == "tensor_product" and p.facts["mode"] == mode and p.facts["n1d"] == 8
    assert classify(E.div_components()).kernel_id == "generic"
    assert classify(E.matvec_f32()).kernel_id == "generic"


def test_facts():
    p = classify(E.grad())
    assert dict(p.facts) == {"n_outer": 3, "n_i": 35, "n_j": 35} and p.long_index == "e"
    p = classify(E.lift_fe())
    assert dict(p.facts) == {"n_outer": 4, "n_i": 35, "n_j": 15}
    assert p.perm == (0, 1, 2)


def test_renaming_and_operand_order_invariance():
    # canonical-form spelling from the reference DB: acd,cbe,de->adb  
    e = f.einsum("acd,cbe,de->adb", f.array("P", (3, 3, "N")), f.array("Q", (3, 35, 35)),
                 f.array("R", ("N", 35)))
    p = classify(e)
    assert p.kernel_id == "grad" and p.long_index == "d"
    # operands permuted: u, J, D
    e = f.einsum("ej,xre,rij->xei", f.array("u", ("E", 35)), f.array("J", (3, 3, "E")),
                 f.array("D", (3, 35, 35)))
    p = classify(e)
    assert p.kernel_id == "grad" and p.perm == (1, 2, 0)


def test_layout_changes_fall_back_to_generic():
    # output transposed: not the layout the grad kernel writes
    e = f.einsum("xre,rij,ej->exi", f.array("J", (3, 3, "E")), f.array("D", (3, 35, 35)),
                 f.array("u", ("E", 35)))
    assert classify(e).kernel_id == "generic"
    # mixed dtypes
    e = f.einsum("xre,rij,ej->xei", f.array("J", (3, 3, "E")), f.array("D", (3, 35, 35), "float32"),
                 f.array("u", ("E", 35)))
    assert classify(e).kernel_id == "generic"
    # J differs between rows
    e = f.batched_einsum("xre,rij,ej->xei", [
        [f.array(f"J{k}", (3, 3, "E")), f.array("D", (3, 35, 35)), f.array(f"u{k}", ("E", 35))]
        for k in range(2)])
    assert classify(e).kernel_id == "generic"
    # no symbolic axis
    e = f.einsum("xre,rij,ej->xei", f.array("J", (3, 3, 10)), f.array("D", (3, 35, 35)),
                 f.array("u", (10, 35)))
    assert classify(e).kernel_id == "generic"


def test_match_subscripts_bijection():
    assert match_subscripts(E.div(), "xre,rij,ej->xei") is None
    perm, imap = match_subscripts(E.div(), "xre,rij,xej->ei")
    assert perm == (0, 1, 2) and imap["e"] == "e"
    e = f.einsum("ij,ij->i", f.array("A", (3, 4)), f.array("B", (3, 4)))
    assert match_subscripts(e, "ij,ik->i") is None


def test_program_params_are_functional():
    prog = generate_cuda(E.div())
    p2 = prog.with_params(variant=2, tile_e=32)
    assert dict(prog.params) == {} and dict(p2.params) == {"variant": 2, "tile_e": 32}
    assert p2.kernel_id == "div" and np.dtype("float64") in set(E.div().arg_to_dtype.values())
