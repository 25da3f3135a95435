"""wave_3d_p4 operator (one call = div + grad + 4-field lift) against the oracle.  GPU only."""

import numpy as np
import pytest

from feinsum_b200 import wave3d
from oracle import np_oracle

pytestmark = pytest.mark.gpu


def _inputs(n, dtype, seed=0):
    rng = np.random.default_rng(seed)
    ins, _ = wave3d.shapes(n)
    return {k: rng.random(s).astype(dtype) for k, s in sorted(ins.items())}


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n", [1, 2, 17, 1000, 10007])
def test_wave3d_matches_oracle(cq, n, dtype):
    import torch

    host = _inputs(n, dtype)
    dev = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in host.items()}
    evt, outs = wave3d.Wave3DExecutor(cq, dtype)(cq, **dev)
    evt.wait()
    got = {k: v.cpu().numpy() for k, v in outs.items()}
    es = wave3d.wave3d_einsums(dtype)
    ref_fn = np_oracle.reference_outputs_fp64 if dtype == "float32" else np_oracle.reference_outputs
    ref_div = ref_fn(es["div"], {k: host[k] for k in ("J", "D", "v")})
    ref_grad = ref_fn(es["grad"], {k: host[k] for k in ("J", "D", "u")})
    ref_lift = ref_fn(es["lift"], {k: host[k] for k in ("L", "Jface", "F_0", "F_1", "F_2", "F_3")})
    np_oracle.assert_matches({"_fe_out": got["div_out"]}, ref_div, north_star=True)
    np_oracle.assert_matches({"_fe_out": got["grad_out"]}, ref_grad, north_star=True)
    lift_names = ["_fe_out", "_fe_out_0", "_fe_out_1", "_fe_out_2"]
    np_oracle.assert_matches({ln: got[f"lift_{k}"] for k, ln in enumerate(lift_names)}, ref_lift, north_star=True)


def test_wave3d_preallocated_outputs_and_errors(cq):
    import torch

    host = _inputs(64, "float64")
    dev = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in host.items()}
    _, out_shapes = wave3d.shapes(64)
    pre = {k: torch.zeros(s, dtype=torch.float64, device=cq.torch_device) for k, s in out_shapes.items()}
    ex = wave3d.Wave3DExecutor(cq)
    evt, outs = ex(cq, **dev, **pre)
    evt.wait()
    assert all(outs[k] is pre[k] for k in pre)
    assert float(outs["div_out"].abs().sum()) > 0
    with pytest.raises(TypeError):
        ex(cq, **{k: v for k, v in dev.items() if k != "Jface"})
    with pytest.raises(ValueError):
        ex(cq, **{**dev, "u": dev["u"][:32]})
    with pytest.raises(TypeError):
        ex(cq, bogus=dev["u"], **dev)
