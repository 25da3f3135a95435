"""Lowering of BatchedEinsums to sm_100a kernels (replaces ``feinsum.codegen``)."""

from feinsum_b200.codegen.cuda import (
    CudaExecutor,
    CudaProgram,
    KernelPlan,
    classify,
    generate_cuda,
    match_subscripts,
)

__all__ = [
    "CudaExecutor",
    "CudaProgram",
    "KernelPlan",
    "classify",
    "generate_cuda",
    "match_subscripts",
]
