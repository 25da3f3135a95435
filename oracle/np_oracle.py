"""
numpy restatement of the reference's validation path.  TEST INFRASTRUCTURE.

Follows, line by line:

* input generation  -- reference ``src/feinsum/measure.py:63-108``
  (``default_rng(seed)``; floats ``rng.random(shape, dtype)`` in [0, 1);
  ints ``rng.integers(-100, 100)``; complex = re + 1j*im).  The reference
  iterates an ``immutables.Map`` whose order depends on string hashing
  (``measure.py:101-108``); here operands are drawn in **sorted-name order**
  so that inputs are reproducible across processes and ranks.
* expected outputs  -- reference ``measure.py:145-159``: one
  ``np.einsum(einsum.get_subscripts(), *row, optimize="optimal")`` per row,
  named ``_fe_out``, ``_fe_out_0``, ...
* tolerances        -- reference ``measure.py:178-185`` (1e-10 fp64 / 1e-6
  fp32, both abs and rel) and the tighter north-star ones (rtol 1e-12 / 1e-5).
"""

from __future__ import annotations

from typing import Any

import numpy as np

REFERENCE_TOL = {np.dtype("float64"): 1e-10, np.dtype("float32"): 1e-6}
NORTH_STAR_RTOL = {np.dtype("float64"): 1e-12, np.dtype("float32"): 1e-5}


def _is_int(x: Any) -> bool:
    return isinstance(x, (int, np.integer))


def concrete_shape(shape: tuple[Any, ...], long_dim_length: int) -> tuple[int, ...]:
    """Replace symbolic extents by ``long_dim_length`` (``measure.py:91-97``)."""
    return tuple(int(d) if _is_int(d) else int(long_dim_length) for d in shape)


def random_array(
    rng: np.random.Generator, dtype: np.dtype[Any], shape: tuple[int, ...]
) -> np.ndarray:
    """reference ``measure.py:63-77``."""
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        real = np.empty(0, dtype).real.dtype
        return (
            rng.random(size=shape, dtype=real)
            + dtype.type(1j) * rng.random(size=shape, dtype=real)
        ).astype(dtype)
    if dtype.kind == "i":
        return rng.integers(low=-100, high=100, size=shape, dtype=dtype)
    return rng.random(size=shape, dtype=dtype)


def generate_input_arrays(
    einsum: Any, long_dim_length: int, np_seed: int = 0
) -> dict[str, np.ndarray]:
    """Host inputs for every distinct operand of *einsum* (sorted by name)."""
    rng = np.random.default_rng(np_seed)
    out: dict[str, np.ndarray] = {}
    for name in sorted(einsum.arg_to_dtype):
        shape = concrete_shape(einsum.arg_to_shape[name], long_dim_length)
        out[name] = random_array(rng, einsum.arg_to_dtype[name], shape)
    return out


def output_names(einsum: Any) -> list[str]:
    """reference ``measure.py:147``."""
    return ["_fe_out", *[f"_fe_out_{i}" for i in range(einsum.b - 1)]]


def reference_outputs(
    einsum: Any, arrays: dict[str, np.ndarray]
) -> dict[str, np.ndarray]:
    """The reference's acceptance oracle (``measure.py:149-159``)."""
    subscripts = einsum.get_subscripts()
    return {
        name: np.einsum(
            subscripts, *[arrays[arg.name] for arg in row], optimize="optimal"
        )
        for name, row in zip(output_names(einsum), einsum.args)
    }


def reference_outputs_fp64(
    einsum: Any, arrays: dict[str, np.ndarray]
) -> dict[str, np.ndarray]:
    """Same, evaluated in float64 and cast back: the fp32 comparisons use this
    so that the oracle's own rounding does not eat the tolerance."""
    subscripts = einsum.get_subscripts()
    outs = {}
    for name, row in zip(output_names(einsum), einsum.args):
        res_dtype = np.result_type(*[arrays[a.name].dtype for a in row])
        wide = [arrays[a.name].astype(np.float64) for a in row]
        outs[name] = np.einsum(subscripts, *wide, optimize="optimal").astype(res_dtype)
    return outs


def assert_matches(
    got: dict[str, np.ndarray],
    ref: dict[str, np.ndarray],
    *,
    north_star: bool = True,
) -> None:
    """reference ``measure.py:167-192`` with selectable tolerance set.

    ``north_star=True``: the drop-in contract's rtol 1e-12 (fp64) / 1e-5 (fp32),
    element by element.  The only absolute slack is 1e-3 of that, scaled by the
    largest reference magnitude -- it exists for exact zeros, not as headroom
    (inputs are in [0,1): the outputs are sums of <= ~10^2 positive products, no
    cancellation)."""
    if set(got) != set(ref):
        raise RuntimeError(f"Output names mismatch: {sorted(got)} vs {sorted(ref)}")
    for name in sorted(ref):
        r, g = ref[name], got[name]
        if r.dtype != g.dtype:
            raise RuntimeError(f"dtype mismatch for output '{name}'")
        if r.shape != g.shape:
            raise RuntimeError(f"shape mismatch for output '{name}'")
        real = np.empty(0, r.dtype).real.dtype
        if north_star:
            rtol = NORTH_STAR_RTOL[np.dtype(real)]
            scale = float(np.max(np.abs(r))) if r.size else 1.0
            np.testing.assert_allclose(g, r, rtol=rtol, atol=1e-3 * rtol * scale, err_msg=name)
        else:
            tol = REFERENCE_TOL[np.dtype(real)]
            np.testing.assert_allclose(g, r, rtol=tol, atol=tol, err_msg=name)


# known answers pinned by the reference's tests / shipped database ----------
# (reference test/test_loopy_utils.py:267-271; data/transform_archive_v5.sqlite
#  giga_op_info column)
KNOWN_FLOPS_PER_ELEMENT = {
    "grad_p4_trivial": 33075,
    "grad_p4": 7980,
    "div_p4": 7980,
    "lift_p4_b4": 17040,
    "tensor_product_p7": 8192,
}
KNOWN_BYTES_PER_ELEMENT_FP64 = {
    "grad_p4": 1192,
    "div_p4": 1192,
    "lift_p4_b4": 3072,
    "tensor_product_p7": 8192,
}
