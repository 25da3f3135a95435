"""Launch-configuration space of the face-mass lift ifj,fe,fej->ei (reference tuning/impls/ifj_fe_fej_to_ei*.py).

Tunables of the hand-written kernels (what the reference's loopy schedule parameters
n_e_per_wg / nwork_items_per_e / i_tiles / j_tiles were for its generated code):

* ``warps``   -- warps of the persistent CTA (one CTA per SM); every warp owns a shared-memory
  slot + output stage, so the legal maximum is set by the 227 KB of an SM
  (``FNSM_E_BAD_CONFIG`` -> InvalidParameterError above it);
* ``variant`` -- 1 = mma.sync tensor path (fp64 DMMA / fp32 3xTF32), 2 = simt fallback, 3 = tcgen05 3xTF32 with
  TMEM accumulators (fp32 only; no further tunables: ``warps`` is ignored).  fp64 einsums tune over variant 1,
  fp32 einsums over 1..3.
"""

from typing import Any

from feinsum_b200.codegen.cuda import CudaProgram
from feinsum_b200.tuning import IntParameter, transform_param

def _variant_space(ensm: Any) -> IntParameter:
    import numpy as np

    fp32 = all(np.dtype(dt) == np.float32 for dt in ensm.arg_to_dtype.values())
    return IntParameter(1, 3) if fp32 else IntParameter(1, 1)


KERNEL_ID = "lift_fe"


@transform_param("warps", lambda ensm: IntParameter(8, 16))
@transform_param("variant", _variant_space)
def transform(program: CudaProgram, warps: int, variant: int = 1, insn_match: Any | None = None,
              kernel_name: str | None = None) -> CudaProgram:
    if program.kernel_id != KERNEL_ID:
        raise ValueError(f"expected a '{KERNEL_ID}' einsum, got '{program.kernel_id}'")
    return program.with_params(threads=32 * int(warps), variant=int(variant))
