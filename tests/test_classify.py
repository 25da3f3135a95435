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
    assert classify(E.div_components()).kernel_id == "opmat_se"
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


def test_shapes_the_opmat_kernels_cannot_hold_go_to_generic():
    """extents without a tensor instantiation that the simt fallback rejects (n_outer > 4 for
    grad/div; operator + element tile beyond the 227 KB of shared memory) must reach the generic kernel
    instead of failing at launch -- the reference runs any BatchedEinsum."""
    from feinsum_b200.codegen.cuda import opmat_kernel_available

    assert classify(E.grad(ndim=5, ndof=6)).kernel_id == "generic"
    assert classify(E.div(ndim=5, ndof=6)).kernel_id == "generic"
    assert classify(E.grad(ndim=4, ndof=6)).kernel_id == "grad"
    # p = 7 tets: D(3,120,120) fp64 = 345 KB
    assert classify(E.grad(ndof=120)).kernel_id == "generic"
    assert classify(E.div(ndof=120)).kernel_id == "generic"
    assert classify(E.grad(ndof=120, dtype="float32")).kernel_id == "grad"      # 173 KB + tile fits
    assert classify(E.lift_ef(nface=6, nvol=125, nfd=25)).kernel_id == "lift_ef"  # 150 KB + tile
    assert classify(E.lift_ef(nface=6, nvol=216, nfd=36)).kernel_id == "generic"  # 373 KB
    # every shape with a tensor-core kernel stays where it was
    for nd, nfd in ((4, 3), (10, 6), (20, 10), (35, 15)):
        assert classify(E.grad(ndof=nd)).kernel_id == "grad"
        assert classify(E.lift_fe(nvol=nd, nfd=nfd)).kernel_id == "lift_fe"
    assert opmat_kernel_available("div", np.dtype("float64"), 3, 56, 56)          # p = 5: simt, 75 KB
    assert not opmat_kernel_available("div", np.dtype("float64"), 3, 120, 120)


def test_loopy_only_entry_points_say_what_to_use_instead():
    """reference src/feinsum/__init__.py:3-6,19-24: names resolve, calls raise NotImplementedError."""
    import pytest

    for name in ("generate_loopy", "generate_loopy_with_opt_einsum_schedule", "get_a_matched_einsum",
                 "get_call_ids", "identify_as_einsum", "match_t_unit_to_einsum"):
        fn = getattr(f, name)
        with pytest.raises(NotImplementedError, match="B200 backend"):
            fn(E.grad())
    with pytest.raises(AttributeError):
        f.no_such_name  # noqa: B018


def test_wave3d_program_host_spec():
    from feinsum_b200 import wave3d
    from feinsum_b200.einsum import SizeParam

    spec = wave3d.Wave3DProgram("float32").host_spec()
    assert set(spec.in_shapes) == set(wave3d.INPUTS) and set(spec.out_shapes) == set(wave3d.OUTPUTS)
    ins, outs = wave3d.shapes(7)
    for n, s in {**spec.in_shapes, **spec.out_shapes}.items():
        conc = tuple(7 if isinstance(d, SizeParam) else d for d in s)
        assert conc == {**ins, **outs}[n]
    assert spec.in_shapes["D"] == (3, 35, 35) and spec.in_dtypes["J"] == np.dtype("float32")


def test_shared_operator_family_se():
    """se,sij,ej->ei (reference test/test_codegen.py:34-88): rows with their own J and u, one operator."""
    p = classify(E.div_components())
    assert p.kernel_id == "opmat_se" and dict(p.facts) == {"n_s": 3, "n_i": 35, "n_j": 35, "es": 0}
    # J(E,3): the layout of reference examples/dg_wave_div.py:14 / test/test_feinsum.py:42
    e = f.batched_einsum("es,sij,ej->ei", [[f.array(f"J{c}", ("E", 3)), f.array("R", (3, 35, 35)),
                                            f.array(f"u{c}", ("E", 35))] for c in "xyz"])
    assert dict(classify(e).facts)["es"] == 1
    assert classify(E.face_mass_se()).kernel_id == "opmat_se"
    # renamed, operands permuted
    e = f.batched_einsum("aq,ra,rpq->ap", [[f.array(f"w{k}", ("N", 20)), f.array(f"G{k}", (3, "N")),
                                            f.array("Op", (3, 20, 20))] for k in range(2)])
    p = classify(e)
    assert p.kernel_id == "opmat_se" and p.perm == (1, 2, 0)
    # operator differs between rows / no compiled shape / fp32 -> generic
    e = f.batched_einsum("se,sij,ej->ei", [[f.array("J", (3, "E")), f.array(f"R{k}", (3, 35, 35)),
                                            f.array(f"u{k}", ("E", 35))] for k in range(2)])
    assert classify(e).kernel_id == "generic"
    e = f.einsum("se,sij,ej->ei", f.array("J", (5, "E")), f.array("R", (5, 7, 7)), f.array("u", ("E", 7)))
    assert classify(e).kernel_id == "generic"
    assert classify(E.div_components("float32")).kernel_id == "generic"
