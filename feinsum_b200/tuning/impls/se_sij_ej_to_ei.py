"""Launch-configuration space of the shared-operator family se,sij,ej->ei (reference
tuning/impls/re_rij_ej_to_ei*.py, re_rji_ej_to_ei_3d_cross_product_v0.py; test einsums
test/test_codegen.py:34-88).

* ``ctas_per_sm`` -- resident 8-warp CTAs per SM of the warp-per-chunk DMMA kernel (0 = as many as fit);
  the p = 4 volume shape (S = 3, 35 dofs) runs the persistent divergence kernel (NX = 1) and ignores it.
"""

from typing import Any

from feinsum_b200.codegen.cuda import CudaProgram
from feinsum_b200.tuning import IntParameter, transform_param

KERNEL_ID = "opmat_se"


@transform_param("ctas_per_sm", lambda ensm: IntParameter(0, 4))
def transform(program: CudaProgram, ctas_per_sm: int = 0, insn_match: Any | None = None,
              kernel_name: str | None = None) -> CudaProgram:
    if program.kernel_id != KERNEL_ID:
        raise ValueError(f"expected a '{KERNEL_ID}' einsum, got '{program.kernel_id}'")
    return program.with_params(ctas_per_sm=int(ctas_per_sm))
