"""Parity of the CUDA path (through the C ABI) against the oracle.  GPU only."""

import json
import os

import numpy as np
import pytest

import feinsum_b200 as f
from feinsum_b200.codegen import generate_cuda
from oracle import np_oracle
from tests import einsums as E
from tests.test_oracle import CASES, load_case

pytestmark = pytest.mark.gpu


def run(einsum, host_inputs, cq, **params):
    import torch

    prog = generate_cuda(einsum)
    if params:
        prog = prog.with_params(**params)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cq.torch_device) for k, v in host_inputs.items()}
    evt, outs = prog.executor(cq)(cq, **dev)
    evt.wait()
    return {k: v.cpu().numpy() for k, v in outs.items()}


def check(einsum, n, cq, seed=0, **params):
    ins = np_oracle.generate_input_arrays(einsum, n, seed)
    got = run(einsum, ins, cq, **params)
    fp32 = any(np.dtype(d) == np.float32 for d in einsum.arg_to_dtype.values())
    ref = (np_oracle.reference_outputs_fp64 if fp32 else np_oracle.reference_outputs)(einsum, ins)
    np_oracle.assert_matches(got, ref, north_star=True)


@pytest.mark.parametrize("name", CASES)
def test_golden_fixtures(golden_dir, cq, name):
    e, ins, outs, _ = load_case(golden_dir, name)
    got = run(e, ins, cq)
    np_oracle.assert_matches(got, outs, north_star=False)
    np_oracle.assert_matches(got, outs, north_star=True)


SIZES = [1, 2, 15, 16, 17, 100, 1001, 10007]


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n", SIZES)
def test_div_fp64(cq, n, variant):
    check(E.div(), n, cq, variant=variant)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n", SIZES)
def test_grad_fp64(cq, n, variant):
    check(E.grad(), n, cq, variant=variant)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("builder", [E.lift_ef, E.lift_fe])
def test_lift_fp64(cq, n, variant, builder):
    check(builder(), n, cq, variant=variant)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n", [1, 4, 17, 1000, 1001, 10008])
@pytest.mark.parametrize("builder", [E.grad, E.div, E.lift_ef, E.lift_fe])
def test_opmat_fp32(cq, n, builder, variant):
    # variant 1 = 3xTF32 tensor path (TMA when n % 4 == 0, plain loads otherwise), 2 = simt
    check(builder(dtype="float32"), n, cq, variant=variant)


@pytest.mark.parametrize("variant", [0, 3])
@pytest.mark.parametrize("n", [4, 128, 132, 1000, 10008, 75776 + 4, 200000])
@pytest.mark.parametrize("builder", [E.grad, E.div, E.lift_ef, E.lift_fe])
def test_opmat_fp32_tcgen05(cq, n, builder, variant):
    # variant 3 = tcgen05 3xTF32 (tiles of 128 elements, two groups per CTA, persistent over 148 SMs:
    # 75 776 elements = two full rounds of tiles, 200 000 = several tiles per group with a ragged tail);
    # variant 0 (auto) must pick it for these sizes (n % 4 == 0)
    check(builder(dtype="float32"), n, cq, variant=variant)


@pytest.mark.parametrize("variant", [0, 3])
@pytest.mark.parametrize("n", [4, 1000, 10008, 200000])
@pytest.mark.parametrize("order", [(4, 3), (10, 6), (20, 10)])
def test_opmat_fp32_tcgen05_other_orders(cq, n, order, variant):
    # tets of order p = 1..3: the tcgen05 kernels are compiled for (volume dofs, face dofs) =
    # (4, 3), (10, 6), (20, 10), (35, 15)
    nd, nfd = order
    check(E.grad(dtype="float32", ndof=nd), n, cq, variant=variant)
    check(E.div(dtype="float32", ndof=nd), n, cq, variant=variant)
    check(E.lift_fe(dtype="float32", nvol=nd, nfd=nfd), n, cq, variant=variant)
    check(E.lift_ef(dtype="float32", nvol=nd, nfd=nfd), n, cq, variant=variant)


@pytest.mark.parametrize("n", [1, 15, 16, 17, 1001, 10007, 100000])
@pytest.mark.parametrize("order", [(4, 3), (10, 6), (20, 10)])
def test_fp32_generic_tensor_kernel_lower_orders(cq, n, order):
    # variant 1 at p = 1..3 = the generic 3xTF32 mma.sync kernel (opmat_tf32_gen.cuh; plain loads, any n);
    # it is also what auto falls back to when n % 4 != 0 rules the tcgen05 kernels out
    nd, nfd = order
    for e in (E.grad(dtype="float32", ndof=nd), E.div(dtype="float32", ndof=nd),
              E.lift_fe(dtype="float32", nvol=nd, nfd=nfd), E.lift_ef(dtype="float32", nvol=nd, nfd=nfd, b=3)):
        check(e, n, cq, variant=1)
    ins = np_oracle.generate_input_arrays(E.div(dtype="float32", ndof=nd), 4000, 5)
    got = run(E.div(dtype="float32", ndof=nd), ins, cq, variant=1)
    ref = np_oracle.reference_outputs_fp64(E.div(dtype="float32", ndof=nd), ins)
    for k in ref:
        assert np.max(np.abs(got[k].astype(np.float64) - ref[k]) / np.abs(ref[k])) < 3e-6


def test_opmat_fp32_orders_without_tensor_kernel(cq):
    # no compiled instantiation (2-D triangles, odd sizes): auto takes the simt kernel, variant 3 refuses
    e = E.grad(dtype="float32", ndim=2, ndof=15)
    check(e, 1000, cq)
    with pytest.raises(f.CudaBackendError):
        check(e, 1000, cq, variant=3)
    check(E.grad(dtype="float32", ndof=20), 1001, cq)        # n % 4 != 0: the generic kernel takes any n


@pytest.mark.parametrize("variant", [0, 3])
@pytest.mark.parametrize("n", [1, 2, 3, 17, 127, 129, 1001, 10007, 75776 + 5, 200001, 200002])
@pytest.mark.parametrize("builder", [E.grad, E.div, E.lift_ef, E.lift_fe])
def test_opmat_fp32_tcgen05_plain_producer(cq, n, builder, variant):
    """n % 4 != 0: rows are not 16-byte multiples, no tensor map can describe the operands.  Round 1 fell back to
    the mma.sync kernels here (63-77 % of roofline); now the tcgen05 kernels themselves run with a bulk-copy producer
    and bulk stores (TMA = false instantiation), for auto and for an explicit variant 3 alike.  Sizes cover
    a single ragged tile, tile boundaries +-1, two full rounds of tiles + 5 and several tiles per group."""
    check(builder(dtype="float32"), n, cq, variant=variant)


@pytest.mark.parametrize("order", [(4, 3), (10, 6), (20, 10)])
def test_opmat_fp32_tcgen05_plain_producer_other_orders(cq, order):
    nd, nfd = order
    for n in (1, 1001, 10007, 200001):
        check(E.grad(dtype="float32", ndof=nd), n, cq, variant=3)
        check(E.div(dtype="float32", ndof=nd), n, cq, variant=3)
        check(E.lift_fe(dtype="float32", nvol=nd, nfd=nfd), n, cq, variant=3)
        check(E.lift_ef(dtype="float32", nvol=nd, nfd=nfd, b=3), n, cq, variant=3)


@pytest.mark.parametrize("n", [1000, 10008])
@pytest.mark.parametrize("builder", [E.grad, E.div, E.lift_ef, E.lift_fe])
def test_opmat_fp32_tcgen05_plain_producer_forced_on_aligned_operands(cq, n, builder):
    # flags bit 0 forces the plain producer where TMA would qualify: both instantiations must agree bit for bit
    e = builder(dtype="float32")
    ins = np_oracle.generate_input_arrays(e, n, 9)
    a = run(e, ins, cq, variant=3)
    b = run(e, ins, cq, variant=3, flags=1)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("variant", [1, 3])
def test_fp32_tensor_path_accuracy_margin(cq, variant):
    """3xTF32 must sit well inside the fp32 north-star tolerance (1e-5): measure the actual
    worst relative error of the tensor paths against the fp64 oracle on 10 000 elements."""
    for builder in (E.grad, E.div, E.lift_fe):
        e = builder(dtype="float32")
        ins = np_oracle.generate_input_arrays(e, 10000, 3)
        got = run(e, ins, cq, variant=variant)
        ref = np_oracle.reference_outputs_fp64(e, ins)
        for k in ref:
            rel = np.max(np.abs(got[k].astype(np.float64) - ref[k]) / np.abs(ref[k]))
            assert rel < 3e-6, (builder.__name__, k, rel)


@pytest.mark.parametrize("n", [16, 1000, 1024, 2049])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("builder", [E.grad, E.div, E.lift_fe, E.lift_ef])
def test_misaligned_operands_take_the_plain_path(cq, builder, dtype, n):
    """Operands that are views at an odd offset into a larger allocation (bases not 16-byte aligned) rule out every
    TMA tensor map: auto takes the TMA = false instantiations (1-D bulk copies started at the address rounded down to
    16 bytes, slabs shifted by the misalignment in shared memory; bulk stores of the aligned middle of a block + its
    ragged ends) and still matches the oracle; outputs are views at an odd offset as well, and nothing outside them
    is written.  Sizes: one chunk / tile, a ragged tail, whole chunks up to the end of the arrays (the last chunk must
    not be read past its end), one element more than that."""
    import torch

    e = builder(dtype=dtype)
    ins = np_oracle.generate_input_arrays(e, n, 11)
    tdt = torch.float64 if dtype == "float64" else torch.float32
    dev = {}
    for pos, (k, v) in enumerate(ins.items()):
        off = 1 + (pos % 3 if dtype == "float32" else 0)   # fp32: 4, 8 and 12 bytes off a 16-byte boundary
        flat = torch.zeros(v.size + 8, dtype=tdt, device=cq.torch_device)
        view = flat[off:off + v.size].view(v.shape)        # base + 1..3 elements
        view.copy_(torch.from_numpy(np.ascontiguousarray(v)))
        assert view.data_ptr() % 16 != 0 and view.is_contiguous()
        dev[k] = view
    out_shape = tuple(int(d) if isinstance(d, (int, np.integer)) else n for d in e.shape)
    flats = {}
    for name in e.output_names:
        flat = torch.full((int(np.prod(out_shape)) + 8,), -7.0, dtype=tdt, device=cq.torch_device)
        flats[name] = flat
        dev[name] = flat[1:1 + int(np.prod(out_shape))].view(out_shape)
    evt, outs = generate_cuda(e).executor(cq)(cq, **dev)
    evt.wait()
    got = {k: v.cpu().numpy() for k, v in outs.items()}
    ref = (np_oracle.reference_outputs_fp64 if dtype == "float32" else np_oracle.reference_outputs)(e, ins)
    np_oracle.assert_matches(got, ref, north_star=True)
    for name, flat in flats.items():                       # canaries around the output views
        host = flat.cpu().numpy()
        assert host[0] == -7.0 and np.all(host[1 + int(np.prod(out_shape)):] == -7.0), name


@pytest.mark.parametrize("n", [1, 100])
def test_grad_batched(cq, n):
    check(E.grad_batched(3), n, cq)
    check(E.grad_batched(9), n, cq)      # more rows than one launch group


@pytest.mark.parametrize("ndof,ndim", [(4, 3), (10, 3), (20, 3), (6, 2), (15, 2)])
def test_other_orders_take_the_simt_variant(cq, ndof, ndim):
    check(E.grad(ndim=ndim, ndof=ndof), 101, cq)
    check(E.div(ndim=ndim, ndof=ndof), 101, cq)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 15, 16, 17, 1001, 10007, 100000])
@pytest.mark.parametrize("order", [(4, 3), (10, 6), (20, 10)])
def test_fp64_dmma_lower_orders(cq, n, order, variant):
    # tets p = 1..3: variant 1 = the generic fp64 DMMA kernel (opmat_dmma_gen.cuh; plain loads, any n),
    # variant 2 = simt cross-check, 0 = auto (must take the DMMA kernel and agree)
    nd, nfd = order
    check(E.grad(ndof=nd), n, cq, variant=variant)
    check(E.div(ndof=nd), n, cq, variant=variant)
    check(E.lift_fe(nvol=nd, nfd=nfd), n, cq, variant=variant)
    check(E.lift_ef(nvol=nd, nfd=nfd, b=3), n, cq, variant=variant)


def test_lift_other_orders(cq):
    check(E.lift_ef(b=3, nface=4, nvol=20, nfd=10), 77, cq)
    check(E.lift_fe(b=5, nface=3, nvol=10, nfd=4), 77, cq)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("n1d", [2, 3, 4, 7, 8])
def test_tensor_product(cq, dtype, mode, n1d):
    for n in (1, 33, 1000):
        check(E.tensor_product(mode, n1d, dtype), n, cq)


def test_tensor_product_persistent_grid(cq):
    check(E.tensor_product(0, 8), 5000, cq, ctas_per_sm=2)


@pytest.mark.parametrize("builder", [E.div_components, E.face_mass_se, E.matvec_f32])
def test_reference_test_einsums_generic(cq, builder):
    # reference test/test_codegen.py:34-97, test/test_measure.py:33-52 (E = 300)
    check(builder(), 300, cq)


def test_generic_misc(cq):
    check(E.matvec_f32(long=True), 1000, cq)
    e = f.einsum("iij,j->i", f.array("A", (5, 5, 7)), f.array("x", 7))
    check(e, 1, cq)
    e = f.einsum("ijk->ij", f.array("P", ("I", 72, 4)))   # reference tuning test einsum
    check(e, 100, cq)
    e = f.einsum("eij,ej->", f.array("A", ("E", 3, 4)), f.array("x", ("E", 4)))  # scalar output
    check(e, 50, cq)


def test_trivial_vs_optimal_schedule_agree(cq):
    # reference test/test_codegen.py:123-165: generic kernel (trivial schedule) vs
    # specialised kernel (hoisted) on the same inputs
    import torch
    from feinsum_b200.codegen.cuda import CudaProgram, KernelPlan

    e = E.grad()
    ins = np_oracle.generate_input_arrays(e, 5)
    fast = run(e, ins, cq)
    dev = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in ins.items()}
    generic = CudaProgram(e, KernelPlan("generic", (0, 1, 2), long_index="e"))
    evt, outs = generic.executor(cq)(cq, **dev)
    evt.wait()
    np.testing.assert_allclose(outs["_fe_out"].cpu().numpy(), fast["_fe_out"], rtol=1e-12)


def test_preallocated_outputs_and_errors(cq):
    import torch

    e = E.div()
    ins = np_oracle.generate_input_arrays(e, 40)
    dev = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in ins.items()}
    out = torch.zeros((40, 35), dtype=torch.float64, device=cq.torch_device)
    ex = generate_cuda(e).executor(cq)
    evt, outs = ex(cq, _fe_out=out, **dev)
    evt.wait()
    assert outs["_fe_out"] is out
    np_oracle.assert_matches({"_fe_out": out.cpu().numpy()}, np_oracle.reference_outputs(e, ins))
    with pytest.raises(ValueError):
        ex(cq, _fe_out=torch.zeros((41, 35), dtype=torch.float64, device=cq.torch_device), **dev)
    with pytest.raises(TypeError):
        ex(cq, **{**dev, "J": ins["J"]})          # host array is not a device buffer
    with pytest.raises(TypeError):
        ex(cq, **{k: v for k, v in dev.items() if k != "u"})
    with pytest.raises(ValueError):
        ex(cq, **{**dev, "u": dev["u"][:, :39]})  # inconsistent E / non-contiguous
    with pytest.raises(TypeError):
        ex(cq, bogus=out, **dev)
    with pytest.raises(f.InvalidParameterError):
        generate_cuda(e).with_params(variant=7).executor(cq)(cq, **dev)


def test_empty_batch(cq):
    import torch

    e = E.div()
    dev = {
        "J": torch.zeros((3, 3, 0), dtype=torch.float64, device=cq.torch_device),
        "D": torch.zeros((3, 35, 35), dtype=torch.float64, device=cq.torch_device),
        "u": torch.zeros((3, 0, 35), dtype=torch.float64, device=cq.torch_device),
    }
    evt, outs = generate_cuda(e).executor(cq)(cq, **dev)
    evt.wait()
    assert tuple(outs["_fe_out"].shape) == (0, 35)


def test_unaligned_views_take_the_plain_load_path(cq):
    # element count odd AND base pointers offset by 8 bytes: TMA path must not be taken
    import torch

    e = E.div()
    n = 1001
    ins = np_oracle.generate_input_arrays(e, n)
    dev = {}
    for k, v in ins.items():
        buf = torch.empty(v.size + 1, dtype=torch.float64, device=cq.torch_device)
        view = buf[1:].view(v.shape)
        view.copy_(torch.from_numpy(v))
        dev[k] = view
    evt, outs = generate_cuda(e).executor(cq)(cq, **dev)
    evt.wait()
    np_oracle.assert_matches({"_fe_out": outs["_fe_out"].cpu().numpy()}, np_oracle.reference_outputs(e, ins))


def test_full_size_div_properties(cq):
    """BASELINE config 2 size (E = 4M): per-element independence lets the oracle
    check slices; linearity in u checks the whole array."""
    import torch

    e = E.div()
    n = 4_000_000
    g = torch.Generator(device=cq.torch_device).manual_seed(1)
    J = torch.rand((3, 3, n), dtype=torch.float64, device=cq.torch_device, generator=g)
    D = torch.rand((3, 35, 35), dtype=torch.float64, device=cq.torch_device, generator=g)
    u = torch.rand((3, n, 35), dtype=torch.float64, device=cq.torch_device, generator=g)
    ex = generate_cuda(e).executor(cq)
    evt, o1 = ex(cq, J=J, D=D, u=u)
    evt.wait()
    out = o1["_fe_out"]
    for sl in (slice(0, 64), slice(1_999_983, 2_000_047), slice(n - 64, n)):
        ins = {"J": J[:, :, sl].cpu().numpy().copy(), "D": D.cpu().numpy(),
               "u": u[:, sl, :].cpu().numpy().copy()}
        np_oracle.assert_matches({"_fe_out": out[sl].cpu().numpy()},
                                 np_oracle.reference_outputs(e, ins))
    # linearity: div(J, D, 2u) == 2 div(J, D, u) exactly (power-of-two scaling)
    evt, o2 = ex(cq, J=J, D=D, u=u * 2.0)
    evt.wait()
    assert torch.equal(o2["_fe_out"], out * 2.0)
    # variant cross-check on the full array (simt vs dmma)
    evt, o3 = generate_cuda(e).with_params(variant=2).executor(cq)(cq, J=J, D=D, u=u)
    evt.wait()
    rel = ((o3["_fe_out"] - out).abs().max() / out.abs().max()).item()
    assert rel < 1e-13


def test_launches_are_counted_and_native(cq):
    from feinsum_b200 import _cabi

    before = _cabi.launch_count()
    check(E.div(), 64, cq)
    assert _cabi.launch_count() > before


def test_shapes_beyond_the_opmat_kernels_run_generic(cq):
    # n_outer = 5 and a 345 KB operator used to fail at launch (UNSUPPORTED / BAD_CONFIG)
    check(E.grad(ndim=5, ndof=6), 50, cq)
    check(E.div(ndim=5, ndof=6), 50, cq)
    check(E.div(ndof=120), 9, cq)


@pytest.mark.parametrize("dtype", ["int32", "int64", "complex64", "complex128"])
def test_generic_integer_and_complex_operands(cq, dtype):
    """The IR accepts ints and complex (reference measure.py:63-77 generates them,
    codegen/loopy.py:258-262 types the result); they run on the generic kernel."""
    e = f.einsum("xre,rij,ej->xei", f.array("J", (2, 2, "E"), dtype), f.array("D", (2, 5, 5), dtype),
                 f.array("u", ("E", 5), dtype))
    ins = np_oracle.generate_input_arrays(e, 37, 2)
    assert ins["u"].dtype == np.dtype(dtype)
    got = run(e, ins, cq)
    ref = np_oracle.reference_outputs(e, ins)
    assert got["_fe_out"].dtype == np.dtype(dtype)
    if np.dtype(dtype).kind == "i":
        assert np.array_equal(got["_fe_out"], ref["_fe_out"])          # integer work: bit-exact
    else:
        np.testing.assert_allclose(got["_fe_out"], ref["_fe_out"],
                                   rtol=1e-5 if dtype == "complex64" else 1e-12)


def test_generic_rejects_mixed_row_dtypes(cq):
    e = f.batched_einsum("ij,j->i", [[f.array("A", ("I", 4), "float64"), f.array("x", 4, "float64")],
                                     [f.array("B", ("I", 4), "float32"), f.array("y", 4, "float32")]])
    with pytest.raises(NotImplementedError):
        generate_cuda(e).executor(cq)


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 100, 1001, 10007, 100000])
@pytest.mark.parametrize("builder", [E.div_components, E.face_mass_se])
def test_shared_operator_family_se(cq, builder, n):
    """se,sij,ej->ei on the FP64 tensor path (reference test/test_codegen.py:34-88): div components = three rows with
    their own J and u on the NX = 1 divergence kernel (TMA for even n, cp.async producer for odd n); face mass = four
    rows sharing J on the warp-per-chunk kernel."""
    assert generate_cuda(builder()).kernel_id == "opmat_se"
    check(builder(), n, cq)


@pytest.mark.parametrize("shape", [(3, 4), (3, 10), (3, 20), (4, 3), (4, 6), (4, 10)])
def test_shared_operator_family_se_other_orders(cq, shape):
    ns, nd = shape
    e = f.batched_einsum("se,sij,ej->ei", [[f.array(f"J{k}", (ns, "E")), f.array("R", (ns, nd, nd)),
                                            f.array(f"u{k}", ("E", nd))] for k in range(3)])
    assert generate_cuda(e).kernel_id == "opmat_se"
    for n in (1, 31, 32, 33, 1001, 20000):
        check(e, n, cq)


def test_shared_operator_family_se_many_rows_and_permuted_operands(cq):
    e = f.batched_einsum("aq,ra,rpq->ap", [[f.array(f"w{k}", ("N", 35)), f.array(f"G{k % 2}", (3, "N")),
                                            f.array("Op", (3, 35, 35))] for k in range(9)])   # 9 rows > one launch group
    assert generate_cuda(e).kernel_id == "opmat_se"
    check(e, 777, cq)
    check(e, 778, cq)


@pytest.mark.parametrize("n", [1, 16, 17, 1001, 10008])
@pytest.mark.parametrize("shape", [(3, 35), (4, 15), (3, 10)])
def test_shared_operator_family_es_layout(cq, shape, n):
    # "es,sij,ej->ei": J(E,S) as in reference examples/dg_wave_div.py:14, test/test_feinsum.py:42
    ns, nd = shape
    e = f.batched_einsum("es,sij,ej->ei", [[f.array(f"J{c}", ("E", ns)), f.array("R", (ns, nd, nd)),
                                            f.array(f"u{c}", ("E", nd))] for c in "xyz"])
    plan = generate_cuda(e).plan
    assert plan.kernel_id == "opmat_se" and plan.facts["es"] == 1
    check(e, n, cq)


@pytest.mark.parametrize("threads", [0, 320, 384])
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 100, 1001, 10007, 10008, 100000])
def test_grad_fp64_second_formulation(cq, n, threads):
    """grad2 (csrc/opmat_grad2.cuh): the divergence kernel's operator tables, (r, nt) column tiles, J applied in
    registers, direct stores -- TMA producer for even n, woven cp.async producer for odd n."""
    params = {"variant": 1, "stages": 2}
    if threads:
        params["threads"] = threads
    check(E.grad(), n, cq, **params)


# ---- "fast start" instantiations of the fp64 DMMA kernels (small launches: griddepcontrol.wait up front + programmatic
# stream serialization, operator staged through shared memory, item-granular lift queue) -- forced on and off at sizes
# on both sides of the automatic switch, TMA and plain producers
@pytest.mark.parametrize("fast_start", [1, 2])
@pytest.mark.parametrize("n", [1, 16, 17, 1000, 1001, 23680, 100001])
@pytest.mark.parametrize("builder", [E.grad, E.div, E.lift_ef, E.lift_fe])
def test_dmma_fast_start(cq, n, builder, fast_start):
    check(builder(), n, cq, variant=1, fast_start=fast_start)


@pytest.mark.parametrize("b", [1, 3, 5, 8])
def test_lift_fast_start_row_counts(cq, b):
    # the item queue divides by the number of rows with a multiplication
    check(E.lift_fe(b=b), 4242, cq, variant=1, fast_start=1)
    check(E.lift_ef(b=b), 4243, cq, variant=1, fast_start=1)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n", [64, 5000, 100000])
def test_dependent_launches_keep_stream_order(cq, n, dtype):
    """lift -> grad -> lift ... back to back on one stream, each kernel consuming what the one before it wrote: the
    fast-start fp64 kernels and the tcgen05 fp32 kernels are launched with programmatic stream serialization and must
    not read before the previous grid has completed."""
    import torch

    lift, grad = E.lift_fe(b=1, dtype=dtype), E.grad(dtype=dtype)
    lin = np_oracle.generate_input_arrays(lift, n, 3)
    gin = np_oracle.generate_input_arrays(grad, n, 4)
    dev_l = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in lin.items()}
    dev_g = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in gin.items()}
    params = dict(variant=1, fast_start=1) if dtype == "float64" else {}
    ex_l = generate_cuda(lift).with_params(**params).executor(cq)
    ex_g = generate_cuda(grad).with_params(**params).executor(cq)
    tdt = getattr(torch, dtype)
    lift_out = torch.zeros((n, 35), dtype=tdt, device=cq.torch_device)
    grad_out = torch.zeros((3, n, 35), dtype=tdt, device=cq.torch_device)
    ref = np_oracle.reference_outputs if dtype == "float64" else np_oracle.reference_outputs_fp64
    ref_l = ref(lift, lin)["_fe_out"]
    ref_g = ref(grad, {**gin, "u": ref_l.astype(dtype)})["_fe_out"]
    for rep in range(20):
        lift_out.zero_()            # a foreign kernel ahead of the chain
        ex_l(cq, **dev_l, _fe_out=lift_out)
        evt, _ = ex_g(cq, **{**dev_g, "u": lift_out}, _fe_out=grad_out)
        evt.wait()
        np_oracle.assert_matches({"_fe_out": grad_out.cpu().numpy()}, {"_fe_out": ref_g}, north_star=True)
