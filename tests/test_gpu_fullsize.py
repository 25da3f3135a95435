"""Parity at the BENCHMARKED sizes (E = 4 000 000 per GPU): every BASELINE config, every kernel family.

The oracle cannot evaluate 4 M elements in seconds, but every output element depends on exactly one
element ``e`` -- so the first, a middle (straddling tile / chunk boundaries) and the last 64-element
slices are checked against ``np_oracle`` exactly as the small cases are, and a size-independent
property covers the whole array: the kernels are linear in the field operand, and scaling it by a
power of two is exact in binary floating point, so ``K(2u) == 2 K(u)`` must hold bit for bit.  (The
one real bug of round 1 -- mbarrier phase aliasing -- only showed at E >= 2 M.)"""

import numpy as np
import pytest

from feinsum_b200 import wave3d
from feinsum_b200.codegen import generate_cuda
from feinsum_b200.einsum import SizeParam
from oracle import np_oracle
from tests import einsums as E

pytestmark = pytest.mark.gpu

N = 4_000_000
SLICES = (slice(0, 64), slice(1_999_983, 2_000_047), slice(N - 64, N))


def _device_inputs(cq, shapes, dtype, seed=1):
    import torch

    g = torch.Generator(device=cq.torch_device).manual_seed(seed)
    tdt = torch.float64 if dtype == "float64" else torch.float32
    return {k: torch.rand(tuple(N if isinstance(d, SizeParam) else int(d) for d in s), dtype=tdt,
                          device=cq.torch_device, generator=g) for k, s in sorted(shapes.items())}


def _cut(sym_shape, t, sl):
    """Slice the element axis of a device tensor and bring it to the host."""
    idx = tuple(sl if isinstance(d, SizeParam) else slice(None) for d in sym_shape)
    return np.ascontiguousarray(t[idx].cpu().numpy())


def _check_slices(e, dev, outs, dtype):
    ref_fn = np_oracle.reference_outputs_fp64 if dtype == "float32" else np_oracle.reference_outputs
    for sl in SLICES:
        ins = {k: _cut(e.arg_to_shape[k], dev[k], sl) for k in e.arg_to_shape}
        got = {k: _cut(e.shape, v, sl) for k, v in outs.items()}
        np_oracle.assert_matches(got, ref_fn(e, ins), north_star=True)


CASES = {
    "grad": (E.grad, "u"), "div": (E.div, "u"),
    "lift_fe": (E.lift_fe, "F_0"), "lift_ef": (E.lift_ef, "v0"),
    "tensor_product": (E.tensor_product, "A"),
}


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_full_size_slices_and_linearity(cq, name, dtype):
    """fp64: the DMMA kernels (TMA path) and the tensor-product kernel; fp32: the tcgen05 / TMEM kernels
    (auto picks them: 4 M % 4 == 0)."""
    import torch

    builder, field = CASES[name]
    e = builder(dtype=dtype)
    dev = _device_inputs(cq, e.arg_to_shape, dtype)
    ex = generate_cuda(e).executor(cq)
    evt, outs = ex(cq, **dev)
    evt.wait()
    _check_slices(e, dev, outs, dtype)
    first = {k: v.clone() for k, v in outs.items()}
    del outs
    # linearity in one field operand, exact under a power-of-two scaling (whole array)
    dev2 = dict(dev)
    dev2[field] = dev[field] * 2.0
    evt, outs2 = ex(cq, **dev2)
    evt.wait()
    name0 = e.output_names[0]
    assert torch.equal(outs2[name0], first[name0] * 2.0)
    for other in e.output_names[1:]:          # rows that do not read `field` are unchanged
        assert torch.equal(outs2[other], first[other])


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_full_size_wave3d(cq, dtype):
    import torch

    prog = wave3d.Wave3DProgram(dtype)
    spec = prog.host_spec()
    dev = _device_inputs(cq, spec.in_shapes, dtype, seed=2)
    evt, outs = prog.executor(cq)(cq, **dev)
    evt.wait()
    es = wave3d.wave3d_einsums(dtype)
    _check_slices(es["div"], dev, {"_fe_out": outs["div_out"]}, dtype)
    _check_slices(es["grad"], dev, {"_fe_out": outs["grad_out"]}, dtype)
    lift_names = ["_fe_out", "_fe_out_0", "_fe_out_1", "_fe_out_2"]
    _check_slices(es["lift"], dev, {ln: outs[f"lift_{k}"] for k, ln in enumerate(lift_names)}, dtype)
    # the operator is the three einsums: whole-array agreement with the stand-alone kernels
    for key, e, names in (("div", es["div"], ["div_out"]), ("grad", es["grad"], ["grad_out"]),
                          ("lift", es["lift"], [f"lift_{k}" for k in range(4)])):
        evt, o = generate_cuda(e).executor(cq)(cq, **{k: dev[k] for k in e.arg_to_shape})
        evt.wait()
        for on, wn in zip(e.output_names, names):
            assert torch.equal(o[on], outs[wn]), (key, wn)
        del o


def test_full_size_alignment_cliff_sizes(cq):
    """E = 4 000 001 / 4 000 002: odd and not-a-multiple-of-4 element counts at full size (plain-load
    producers); first / last slices against the oracle."""
    import torch

    for n, dtype in ((4_000_001, "float64"), (4_000_002, "float32"), (4_000_001, "float32")):
        for builder in (E.grad, E.div, E.lift_fe):
            e = builder(dtype=dtype)
            g = torch.Generator(device=cq.torch_device).manual_seed(3)
            tdt = torch.float64 if dtype == "float64" else torch.float32
            dev = {k: torch.rand(tuple(n if isinstance(d, SizeParam) else int(d) for d in s), dtype=tdt,
                                 device=cq.torch_device, generator=g) for k, s in sorted(e.arg_to_shape.items())}
            evt, outs = generate_cuda(e).executor(cq)(cq, **dev)
            evt.wait()
            ref_fn = np_oracle.reference_outputs_fp64 if dtype == "float32" else np_oracle.reference_outputs
            for sl in (slice(0, 64), slice(n - 70, n)):
                ins = {k: _cut(e.arg_to_shape[k], dev[k], sl) for k in e.arg_to_shape}
                got = {k: _cut(e.shape, v, sl) for k, v in outs.items()}
                np_oracle.assert_matches(got, ref_fn(e, ins), north_star=True)
            del dev, outs
