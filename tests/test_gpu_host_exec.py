"""HostExecutor (numpy in, numpy out; chunked H2D | kernel | D2H pipeline) against the oracle.  GPU only."""

import gc

import numpy as np
import pytest

import feinsum_b200 as f
from feinsum_b200 import wave3d
from feinsum_b200.codegen import generate_cuda
from feinsum_b200.host_exec import HostExecutor, pinned_empty
from oracle import np_oracle
from tests import einsums as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunk", [64, 1000, 4096])
@pytest.mark.parametrize("builder", [E.div, E.grad, E.lift_fe, E.lift_ef, E.tensor_product])
def test_chunked_pipeline_matches_oracle(cq, builder, chunk):
    # n = 3001: odd (no TMA path for the tail chunk), several chunks plus a ragged tail; operands whose
    # element axis is not leading (J(3,3,E), u(3,E,35), out(3,E,35)) go through strided 2-D copies
    e = builder()
    n = 3001
    ins = np_oracle.generate_input_arrays(e, n, 4)
    got = HostExecutor(generate_cuda(e), cq, chunk=chunk)(**ins)
    np_oracle.assert_matches(got, np_oracle.reference_outputs(e, ins), north_star=True)


def test_pinned_inputs_preallocated_outputs_and_reuse(cq):
    e = E.div(dtype="float32")
    n = 2500
    ins = np_oracle.generate_input_arrays(e, n, 5)
    pinned = {}
    for k, v in ins.items():
        pinned[k] = pinned_empty(v.shape, v.dtype)
        pinned[k][...] = v
    out = {"_fe_out": pinned_empty((n, 35), np.float32)}
    hx = HostExecutor(generate_cuda(e), cq, chunk=1024)
    for _ in range(3):                       # buffers are reused across calls
        res = hx(outputs=out, **pinned)
        assert res["_fe_out"] is out["_fe_out"]
        np_oracle.assert_matches(res, np_oracle.reference_outputs_fp64(e, ins), north_star=True)
    assert hx.h2d_bytes == sum(v.nbytes for v in ins.values())
    assert hx.d2h_bytes == out["_fe_out"].nbytes


def test_pinned_buffers_are_released(cq):
    """outputs=None allocates fresh pinned results per call; the numpy view's base chain owns the pinned
    tensor and nothing in the module keeps it alive (round-1 `_KEEPALIVE` leaked every buffer)."""
    import weakref

    import feinsum_b200.host_exec as hx_mod

    assert not hasattr(hx_mod, "_KEEPALIVE")
    a = pinned_empty((1024,), np.float64)
    assert a.base is not None
    owner = weakref.ref(a.base)
    a[:] = 1.0
    del a
    gc.collect()
    assert owner() is None


def test_errors(cq):
    e = E.div()
    ins = np_oracle.generate_input_arrays(e, 10)
    hx = HostExecutor(generate_cuda(e), cq)
    with pytest.raises(TypeError):
        hx(**{k: v for k, v in ins.items() if k != "u"})
    with pytest.raises(TypeError):
        hx(**{**ins, "u": ins["u"].astype(np.float32)})
    with pytest.raises(ValueError):
        hx(**{**ins, "u": ins["u"][:, :9].copy()})
    with pytest.raises(ValueError):
        hx(outputs={"_fe_out": np.zeros((11, 35))}, **ins)
    with pytest.raises(TypeError):
        hx(bogus=ins["u"], **ins)


def test_empty_and_unchunkable(cq):
    e = E.div()
    ins = np_oracle.generate_input_arrays(e, 0)
    assert HostExecutor(generate_cuda(e), cq)(**ins)["_fe_out"].shape == (0, 35)
    # scalar output: the element axis is contracted, nothing to chunk over
    s = f.einsum("eij,ej->", f.array("A", ("E", 3, 4)), f.array("x", ("E", 4)))
    ins = np_oracle.generate_input_arrays(s, 50)
    got = HostExecutor(generate_cuda(s), cq, chunk=8)(**ins)
    np_oracle.assert_matches(got, np_oracle.reference_outputs(s, ins), north_star=True)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_wave3d_through_the_chunk_pipeline(cq, dtype):
    n = 2001
    rng = np.random.default_rng(7)
    ins_shapes, _ = wave3d.shapes(n)
    host = {k: rng.random(s).astype(dtype) for k, s in sorted(ins_shapes.items())}
    got = HostExecutor(wave3d.Wave3DProgram(dtype), cq, chunk=512)(**host)
    es = wave3d.wave3d_einsums(dtype)
    ref_fn = np_oracle.reference_outputs_fp64 if dtype == "float32" else np_oracle.reference_outputs
    np_oracle.assert_matches({"_fe_out": got["div_out"]}, ref_fn(es["div"], {k: host[k] for k in ("J", "D", "v")}))
    np_oracle.assert_matches({"_fe_out": got["grad_out"]}, ref_fn(es["grad"], {k: host[k] for k in ("J", "D", "u")}))
    ref_lift = ref_fn(es["lift"], {k: host[k] for k in ("L", "Jface", "F_0", "F_1", "F_2", "F_3")})
    names = ["_fe_out", "_fe_out_0", "_fe_out_1", "_fe_out_2"]
    np_oracle.assert_matches({ln: got[f"lift_{k}"] for k, ln in enumerate(names)}, ref_lift)
