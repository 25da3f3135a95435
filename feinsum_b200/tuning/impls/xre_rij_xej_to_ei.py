"""Launch-configuration space of the DG divergence xre,rij,xej->ei (reference tuning/impls/xre_rij_xej_to_ei*.py).

Tunables of the hand-written kernels (what the reference's loopy schedule parameters
n_e_per_wg / nwork_items_per_e / i_tiles / j_tiles were for its generated code):

* ``warps``   -- warps of the persistent CTA (one CTA per SM); every warp owns a shared-memory
  slot + output stage, so the legal maximum is set by the 227 KB of an SM
  (``FNSM_E_BAD_CONFIG`` -> InvalidParameterError above it);
* ``variant`` -- 1 = tensor path (fp64 DMMA / fp32 3xTF32, p = 4 shapes), 2 = simt fallback.
"""

from typing import Any

from feinsum_b200.codegen.cuda import CudaProgram
from feinsum_b200.tuning import IntParameter, transform_param

KERNEL_ID = "div"


@transform_param("warps", lambda ensm: IntParameter(8, 16))
@transform_param("variant", lambda ensm: IntParameter(1, 1))
def transform(program: CudaProgram, warps: int, variant: int = 1, insn_match: Any | None = None,
              kernel_name: str | None = None) -> CudaProgram:
    if program.kernel_id != KERNEL_ID:
        raise ValueError(f"expected a '{KERNEL_ID}' einsum, got '{program.kernel_id}'")
    return program.with_params(threads=32 * int(warps), variant=int(variant))
