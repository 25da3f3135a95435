"""Launch-configuration space of the tensor-product sum-factorisation ``eabc,ia->eibc`` (and the two
other contraction modes).  The kernel is a pure streaming kernel; the only tunable is the size of the
persistent grid (0 = one CTA per 256 work items, i.e. not persistent)."""

from typing import Any

from feinsum_b200.codegen.cuda import CudaProgram
from feinsum_b200.tuning import IntParameter, transform_param


@transform_param("ctas_per_sm", lambda ensm: IntParameter(0, 16))
def transform(program: CudaProgram, ctas_per_sm: int, insn_match: Any | None = None,
              kernel_name: str | None = None) -> CudaProgram:
    if program.kernel_id != "tensor_product":
        raise ValueError(f"expected a tensor-product einsum, got '{program.kernel_id}'")
    return program.with_params(ctas_per_sm=int(ctas_per_sm)) if ctas_per_sm else program
