"""Fused three-mode hex derivative (A read once) against the oracle and against the three stand-alone
tensor-product kernels.  GPU only."""

import numpy as np
import pytest

from feinsum_b200 import hexderiv
from feinsum_b200.codegen import generate_cuda
from oracle import np_oracle

pytestmark = pytest.mark.gpu


def _inputs(n, seed=0):
    rng = np.random.default_rng(seed)
    ins, _ = hexderiv.shapes(n)
    return {k: rng.random(s) for k, s in sorted(ins.items())}


@pytest.mark.parametrize("stages", [0, 2, 3, 4])
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 1000, 20011])
def test_hex_deriv_matches_oracle(cq, n, stages):
    import torch

    host = _inputs(n)
    dev = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in host.items()}
    prog = hexderiv.HexDerivProgram(**({"stages": stages} if stages else {}))
    evt, outs = prog.executor(cq)(cq, **dev)
    evt.wait()
    for k, (name, e) in enumerate(hexderiv.hexderiv_einsums().items()):
        ref = np_oracle.reference_outputs(e, {"A": host["A"], f"M{k}": host[f"M{k}"]})
        np_oracle.assert_matches({"_fe_out": outs[name].cpu().numpy()}, ref, north_star=True)


def test_hex_deriv_agrees_with_the_three_tensor_product_kernels(cq):
    import torch

    n = 30000
    host = _inputs(n, 3)
    dev = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in host.items()}
    evt, outs = hexderiv.HexDerivExecutor(cq)(cq, **dev)
    evt.wait()
    for k, (name, e) in enumerate(hexderiv.hexderiv_einsums().items()):
        evt, o = generate_cuda(e).executor(cq)(cq, A=dev["A"], **{f"M{k}": dev[f"M{k}"]})
        evt.wait()
        rel = ((o["_fe_out"] - outs[name]).abs().max() / o["_fe_out"].abs().max()).item()
        assert rel < 1e-14, (name, rel)


def test_hex_deriv_errors_and_preallocated_outputs(cq):
    import torch

    host = _inputs(16)
    dev = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in host.items()}
    ex = hexderiv.HexDerivExecutor(cq)
    pre = {f"d{k}": torch.zeros((16, 8, 8, 8), dtype=torch.float64, device=cq.torch_device) for k in range(3)}
    evt, outs = ex(cq, **dev, **pre)
    evt.wait()
    assert all(outs[k] is pre[k] for k in pre) and float(outs["d2"].abs().sum()) > 0
    with pytest.raises(TypeError):
        ex(cq, **{k: v for k, v in dev.items() if k != "M1"})
    with pytest.raises(ValueError):
        ex(cq, **{**dev, "M0": dev["M0"][:7]})
    with pytest.raises(NotImplementedError):
        hexderiv.HexDerivExecutor(cq, "float32")
    import feinsum_b200 as f

    with pytest.raises(f.InvalidParameterError):
        hexderiv.HexDerivProgram(stages=7).executor(cq)(cq, **dev)


def test_hex_deriv_full_size_slices(cq):
    """E = 4 M (the benchmarked size): first / middle / last slices against the oracle."""
    import torch

    n = 4_000_000
    g = torch.Generator(device=cq.torch_device).manual_seed(5)
    dev = {"A": torch.rand((n, 8, 8, 8), dtype=torch.float64, device=cq.torch_device, generator=g)}
    for k in range(3):
        dev[f"M{k}"] = torch.rand((8, 8), dtype=torch.float64, device=cq.torch_device, generator=g)
    evt, outs = hexderiv.HexDerivExecutor(cq)(cq, **dev)
    evt.wait()
    for sl in (slice(0, 16), slice(1_999_990, 2_000_010), slice(n - 16, n)):
        A = dev["A"][sl].cpu().numpy()
        for k, (name, e) in enumerate(hexderiv.hexderiv_einsums().items()):
            ref = np_oracle.reference_outputs(e, {"A": A, f"M{k}": dev[f"M{k}"].cpu().numpy()})
            np_oracle.assert_matches({"_fe_out": outs[name][sl].cpu().numpy()}, ref, north_star=True)
