"""Launch-configuration space of the DG gradient xre,rij,ej->xei (reference tuning/impls/xre_rij_ej_to_xei.py).

Tunables of the hand-written kernels (what the reference's loopy schedule parameters
n_e_per_wg / nwork_items_per_e / i_tiles / j_tiles were for its generated code):

* ``warps``   -- warps of the persistent CTA (one CTA per SM); every warp owns a shared-memory
  slot + output stage, so the legal maximum is set by the 227 KB of an SM
  (``FNSM_E_BAD_CONFIG`` -> InvalidParameterError above it);
* ``variant`` -- 1 = mma.sync tensor path (fp64 DMMA / fp32 3xTF32), 2 = simt fallback, 3 = tcgen05 3xTF32 with
  TMEM accumulators (fp32 only; no further tunables: ``warps`` is ignored).  fp64 einsums tune over variant 1,
  fp32 einsums over 1..3.

* ``ctas_per_sm`` -- resident CTAs per SM of the warp-per-chunk kernels that run the lower orders (tets p = 1..3:
  ``opmat_dmma_gen.cuh`` / ``opmat_tf32_gen.cuh``) and of the simt kernel; 0 = as many as fit.  The persistent p = 4
  kernels ignore it (one CTA per SM by construction);
* ``tile_e8``  -- simt only (``variant`` 2): elements per CTA tile in units of 8, 0 = library default.
* ``formulation`` -- fp64 p = 4 only: 1 = ``k_grad_dmma`` (own operator table, staged TMA stores), 2 = ``k_grad2_dmma``
  (divergence tables, direct stores; profiles/r02_ab_dmma.md).
"""

from typing import Any

from feinsum_b200.codegen.cuda import CudaProgram
from feinsum_b200.tuning import IntParameter, transform_param

def _variant_space(ensm: Any) -> IntParameter:
    import numpy as np

    fp32 = all(np.dtype(dt) == np.float32 for dt in ensm.arg_to_dtype.values())
    return IntParameter(1, 3) if fp32 else IntParameter(1, 1)


def _is_p4(ensm: Any) -> bool:
    """Tets of order 4 (35 volume dofs): the persistent, hand-tuned kernels; anything else runs the warp-per-chunk /
    simt kernels, whose knobs are the grid and the tile instead of the warp count."""
    import numbers

    return any(isinstance(ext, numbers.Integral) and int(ext) == 35 for ext in ensm.index_to_dim_length.values())


def _fp64(ensm: Any) -> bool:
    import numpy as np

    return all(np.dtype(dt) == np.float64 for dt in ensm.arg_to_dtype.values())


KERNEL_ID = "grad"


@transform_param("warps", lambda ensm: IntParameter(8, 16) if _is_p4(ensm) else IntParameter(8, 8))
@transform_param("variant", _variant_space)
@transform_param("ctas_per_sm", lambda ensm: IntParameter(0, 0) if _is_p4(ensm) else IntParameter(0, 4))
@transform_param("tile_e8", lambda ensm: IntParameter(0, 0) if _is_p4(ensm) else IntParameter(0, 4))
@transform_param("formulation", lambda ensm: IntParameter(1, 2) if _is_p4(ensm) and _fp64(ensm) else IntParameter(1, 1))
def transform(program: CudaProgram, warps: int, variant: int = 1, ctas_per_sm: int = 0, tile_e8: int = 0, formulation: int = 1,
              insn_match: Any | None = None, kernel_name: str | None = None) -> CudaProgram:
    if program.kernel_id != KERNEL_ID:
        raise ValueError(f"expected a '{KERNEL_ID}' einsum, got '{program.kernel_id}'")
    params = {"threads": 32 * int(warps), "variant": int(variant)}
    if int(ctas_per_sm) > 0:
        params["ctas_per_sm"] = int(ctas_per_sm)
    if int(tile_e8) > 0 and int(variant) == 2:
        params["tile_e"] = 8 * int(tile_e8)
    if int(formulation) == 2:
        params["stages"] = 2
    return program.with_params(**params)
