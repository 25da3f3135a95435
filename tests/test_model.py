"""Contraction schedules and the FLOP / byte model, pinned by the reference's
known answers (test/test_loopy_utils.py:267-271, transform_archive_v5.sqlite
giga_op_info)."""

import numpy as np
import pytest

import feinsum_b200 as f
from feinsum_b200 import measure
from feinsum_b200.contraction_schedule import EinsumOperand, IntermediateResult
from oracle.np_oracle import KNOWN_BYTES_PER_ELEMENT_FP64, KNOWN_FLOPS_PER_ELEMENT
from tests import einsums as E

F64 = np.dtype("float64")


def test_trivial_schedule():
    s = f.get_trivial_contraction_schedule(E.grad())
    assert s.nsteps == 1 and s.subscripts == ("xre,rij,ej -> xei",)
    assert s.result_names == ("_fe_out",)
    assert s.arguments == ((EinsumOperand(0), EinsumOperand(1), EinsumOperand(2)),)


@pytest.mark.parametrize(
    "einsum,steps",
    [
        (E.grad(), ("ej,rij->rie", "rie,xre->xei")),
        (E.div(), ("xej,xre->rje", "rje,rij->ei")),
        (E.lift_ef(), ("fej,ef->fje", "fje,fij->ei")),
    ],
)
def test_optimal_paths_match_opt_einsum(einsum, steps):
    # paths recorded in SURVEY.md section 8(a3) (numpy/opt_einsum "optimal")
    s = f.get_opt_einsum_contraction_schedule(einsum)
    assert s.subscripts == steps
    assert s.result_names == ("_fe_tmp", "_fe_out")
    assert isinstance(s.arguments[1][0], IntermediateResult)


def test_optimal_path_agrees_with_numpy_einsum_path_cost():
    # numpy's optimal path must not beat ours in its own flop estimate
    for e in (E.grad(), E.div(), E.lift_ef(), E.tensor_product()):
        ours = sum(measure.get_flops_per_dtype(e, 1000).values())
        triv = sum(
            measure.get_flops_per_dtype(e, 1000, f.get_trivial_contraction_schedule(e)).values()
        )
        assert ours <= triv


def test_known_flops_per_element():
    n = 100_000
    per = lambda e, s=None: measure.get_flops_per_dtype(e, n, s)[F64] / n  # noqa: E731
    g = E.grad()
    # operator-only steps do not exist here: every step carries the element axis
    assert per(g, f.get_trivial_contraction_schedule(g)) == KNOWN_FLOPS_PER_ELEMENT["grad_p4_trivial"]
    assert per(g) == KNOWN_FLOPS_PER_ELEMENT["grad_p4"]
    assert per(E.div()) == KNOWN_FLOPS_PER_ELEMENT["div_p4"]
    assert per(E.lift_ef()) == KNOWN_FLOPS_PER_ELEMENT["lift_p4_b4"]
    assert per(E.lift_fe()) == KNOWN_FLOPS_PER_ELEMENT["lift_p4_b4"]
    assert per(E.tensor_product()) == KNOWN_FLOPS_PER_ELEMENT["tensor_product_p7"]


def test_giga_ops_match_shipped_database():
    # data/transform_archive_v5.sqlite: giga_op_info 0.798 (grad/div) and 1.704 (lift b=4) at E=1e5
    assert measure._get_giga_ops_from_einsum(E.grad())[F64] == pytest.approx(0.798)
    assert measure._get_giga_ops_from_einsum(E.div())[F64] == pytest.approx(0.798)
    assert measure._get_giga_ops_from_einsum(E.lift_fe())[F64] == pytest.approx(1.704)


def test_known_bytes_per_element():
    n = 1_000_000
    const = {"grad_p4": 29400, "div_p4": 29400, "lift_p4_b4": 16800, "tensor_product_p7": 512}
    for name, e in [("grad_p4", E.grad()), ("div_p4", E.div()), ("lift_p4_b4", E.lift_ef()),
                    ("tensor_product_p7", E.tensor_product())]:
        got = (measure.get_footprint_bytes(e, n) - const[name]) / n
        assert got == KNOWN_BYTES_PER_ELEMENT_FP64[name], name
    assert (measure.get_footprint_bytes(E.grad("float32"), n) - 14700) / n == 596


def test_roofline_rates_titan_v_reproduce_baseline_md():
    # feinsum's Titan V roofline 4370 GFLOP/s (grad), 3621 (lift)
    r = measure.get_roofline_flop_rate(E.grad(), "NVIDIA TITAN V")[F64]
    assert r == pytest.approx(4370, rel=2e-3)
    r = measure.get_roofline_flop_rate(E.lift_fe(), "NVIDIA TITAN V")[F64]
    assert r == pytest.approx(3621, rel=2e-3)


def test_unknown_device_roofline():
    with pytest.raises(f.NoDevicePeaksInfoError):
        measure.get_roofline_flop_rate(E.grad(), "No Such Device")
    s = measure._stringify_runtime_comparison_vs_roofline(E.grad(), 1e-3, "No Such Device")
    assert "N/A" in s and "Measured GOps/s" in s
    s = measure._stringify_runtime_comparison_vs_roofline(E.grad(), 1e-3, "NVIDIA B200")
    assert "float64" in s and "N/A" not in s


def test_b200_roofline_report_bounds():
    rep = measure.roofline_report(E.tensor_product(), 5e-3, "NVIDIA B200", 4_000_000)
    assert rep["bound"] == "hbm" and rep["bytes"] == pytest.approx(8192 * 4e6 + 512)
    rep = measure.roofline_report(E.div(), 1e-3, "NVIDIA B200", 4_000_000)
    assert rep["flops"] == 7980 * 4e6 and 0 < rep["roofline_frac"]
