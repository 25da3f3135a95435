"""The oracle checked against the fixtures the REFERENCE produced
(tests/golden/make_golden.py) and against its independent C restatement."""

import json
import os

import numpy as np
import pytest

import feinsum_b200 as f
from oracle import cgen, np_oracle
from tests.einsums import from_spec

CASES = [
    "grad_p4", "div_p4", "lift_p4_b4", "lift_fe_p4_b4", "tensor_product_p7",
    "div_components", "face_mass_se", "matvec_f32", "matvec_f32_long", "diag_access",
    "grad_p4_f32",
    "grad_p2", "div_p3", "lift_fe_p1_b4", "lift_p3_b4", "div_p2_f32",
]


def load_case(golden_dir, name):
    with open(os.path.join(golden_dir, "frontend.json")) as fh:
        spec = json.load(fh)["valid"][name]["spec"]
    data = np.load(os.path.join(golden_dir, f"numeric_{name}.npz"))
    ins = {k[4:]: data[k] for k in data.files if k.startswith("in__")}
    outs = {k[5:]: data[k] for k in data.files if k.startswith("out__")}
    return from_spec(spec), ins, outs, int(data["long_dim_length"])


@pytest.mark.parametrize("name", CASES)
def test_np_oracle_reproduces_reference_outputs(golden_dir, name):
    e, ins, outs, _ = load_case(golden_dir, name)
    got = np_oracle.reference_outputs(e, ins)
    assert set(got) == set(outs)
    for k in outs:
        # same expression, same inputs, same numpy -> bit-identical
        assert got[k].dtype == outs[k].dtype and np.array_equal(got[k], outs[k]), (name, k)


@pytest.mark.parametrize("name", CASES)
def test_c_restatement_matches_reference_outputs(golden_dir, name):
    e, ins, outs, n = load_case(golden_dir, name)
    for sched in (None, f.get_opt_einsum_contraction_schedule(e)):
        got = cgen.CKernel(e, sched)(n, ins)
        np_oracle.assert_matches(got, outs, north_star=False)   # reference tolerances
        np_oracle.assert_matches(got, outs, north_star=True)    # rtol 1e-12 / 1e-5


def test_input_generator_distribution_and_determinism(golden_dir):
    e, ins, _, n = load_case(golden_dir, "grad_p4")
    mine = np_oracle.generate_input_arrays(e, n)
    again = np_oracle.generate_input_arrays(e, n)
    for k in ins:
        assert mine[k].shape == ins[k].shape and mine[k].dtype == ins[k].dtype
        assert np.array_equal(mine[k], again[k])
        assert (mine[k] >= 0).all() and (mine[k] < 1).all()
    # sorted-name draw order: D, J, u -- first draw must equal a fresh rng's first draw
    rng = np.random.default_rng(0)
    assert np.array_equal(mine["D"], rng.random((3, 35, 35)))


def test_fp32_oracle_in_fp64(golden_dir):
    e, ins, outs, _ = load_case(golden_dir, "grad_p4_f32")
    wide = np_oracle.reference_outputs_fp64(e, ins)
    assert wide["_fe_out"].dtype == np.float32
    np.testing.assert_allclose(wide["_fe_out"], outs["_fe_out"], rtol=1e-5)


def test_assert_matches_detects_errors(golden_dir):
    e, ins, outs, _ = load_case(golden_dir, "div_p4")
    bad = {k: v.copy() for k, v in outs.items()}
    bad["_fe_out"][3, 7] *= 1 + 1e-9
    with pytest.raises(AssertionError):
        np_oracle.assert_matches(bad, outs)
    with pytest.raises(RuntimeError):
        np_oracle.assert_matches({}, outs)
