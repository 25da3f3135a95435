"""measure.py on a real device: validation gate, CUDA-event timeit, roofline table."""

import numpy as np
import pytest

import feinsum_b200 as f
from feinsum_b200 import measure
from tests import einsums as E

pytestmark = pytest.mark.gpu

IDENTITY = lambda t_unit, insn_match, kernel_name: t_unit  # noqa: E731  (reference tests' transform)


@pytest.fixture(autouse=True)
def short_timing(monkeypatch):
    monkeypatch.setattr(measure, "N_MIN_SIM_SECS", 0.05)


def test_timeit_reference_style(cq):
    # reference test/test_codegen.py:34-120
    for e in (E.div_components(), E.face_mass_se(), E.grad()):
        t = measure.timeit(e, transform=IDENTITY, cq=cq, long_dim_length=300)
        assert 0 < t < 1


def test_simple_matvec(cq):
    # reference test/test_measure.py:33-52
    measure.timeit(E.matvec_f32(), transform=IDENTITY, cq=cq)
    measure.timeit(E.matvec_f32(long=True), transform=IDENTITY, cq=cq)


def test_pprint_roofline_comparison(cq):
    # reference test/test_measure.py:55-81, with a launch-parameter transform
    s = measure.stringify_comparison_vs_roofline(
        E.grad(), cq=cq,
        transform=lambda t_unit, insn_match, kernel_name: t_unit.with_params(variant=2, tile_e=32),
        long_dim_length=500,
    )
    assert "Measured GOps/s" in s and "float64" in s


def test_validation_gate_raises_on_wrong_kernel(cq):
    from feinsum_b200.codegen.cuda import CudaProgram, KernelPlan

    # force the div kernel onto the grad einsum's data layout -> wrong numbers
    e = E.div()

    def wrong(t_unit, insn_match=None, kernel_name=None):
        return CudaProgram(t_unit.einsum, KernelPlan(
            "lift_fe", (1, 0, 2), t_unit.plan.facts, "e"))

    with pytest.raises((f.TransformValidationError, f.CudaBackendError)):
        measure.validate_batched_einsum_transform(e, cq, wrong)


def test_giga_op_rate_keys(cq):
    r = measure.measure_giga_op_rate(E.div(), transform=IDENTITY, cq=cq, long_dim_length=2000)
    assert set(r) == {np.dtype("float64")} and r[np.dtype("float64")] > 0


def test_measured_peaks(cq):
    from feinsum_b200 import _cabi

    fp64 = _cabi.measure_peak(0)
    assert 5_000 < fp64 < 100_000
    bw = _cabi.measure_peak(2)
    assert 1_000 < bw < 12_000


def test_autotune_record_retrieve_on_device(cq, tmp_path):
    """End to end on the GPU: search the grad launch space for a few points, then retrieve the best
    configuration from the facts database and run it (reference test_tuple_args.py's flow)."""
    import os

    import torch

    from feinsum_b200 import tuning
    from feinsum_b200.codegen import generate_cuda
    from oracle import np_oracle

    db = str(tmp_path / "facts.sqlite")
    mod = os.path.join(tuning._get_impls_path(), "xre_rij_ej_to_xei.py")
    best = f.autotune(E.grad(), mod, cq, db_path=db, long_dim_length=20_000, test_limit=3)
    facts = f.query(E.grad(), cq.device, database=db)
    assert best is not None and 1 <= len(facts) <= 3
    assert all(q.runtime_in_sec > 0 and q.giga_op_rate("float64") > 100 for q in facts)
    transform = f.retrieve(E.grad(), cq.device, database=db)
    prog = transform(generate_cuda(E.grad()), insn_match=None, kernel_name=None)
    ins = np_oracle.generate_input_arrays(E.grad(), 1000)
    dev = {k: torch.from_numpy(v).to(cq.torch_device) for k, v in ins.items()}
    evt, outs = prog.executor(cq)(cq, **dev)
    evt.wait()
    np_oracle.assert_matches({k: v.cpu().numpy() for k, v in outs.items()},
                             np_oracle.reference_outputs(E.grad(), ins), north_star=True)
    # a second search only re-times what is not recorded yet
    n_before = len(facts)
    f.autotune(E.grad(), mod, cq, db_path=db, long_dim_length=20_000, test_limit=1)
    assert len(f.query(E.grad(), cq.device, database=db)) <= n_before + 1


@pytest.mark.parametrize("builder,tol", [(E.grad, 1e-6), (E.lift_fe, 1e-6), (E.lift_ef, 1e-6), (E.div, 2e-6)])
def test_fp32_tensor_path_against_the_reference_gate(cq, builder, tol, monkeypatch):
    """The reference's own acceptance test for fp32 is rtol = atol = 1e-6 against numpy's fp32
    einsum at E = 100 (reference measure.py:178-192).  grad and lift pass it as is; div (105-term
    sums accumulated with tensor-core truncation) needs 2e-6 -- see measure.FP32_VALIDATION_TOL."""
    monkeypatch.setattr(measure, "FP32_VALIDATION_TOL", tol)
    measure.validate_batched_einsum_transform(builder(dtype="float32"), cq, IDENTITY)
