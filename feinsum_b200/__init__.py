"""
feinsum_b200 -- B200-native execution backend for feinsum's batched einsums.

Front-end (``array`` / ``einsum`` / ``batched_einsum`` / ``BatchedEinsum``)
is API-compatible with kaushikcfd/feinsum 2025.3 (reference
``src/feinsum/__init__.py:37-68``); kernels are hand-written sm_100a CUDA
reached through a C ABI (``include/fnsm_b200.h``).  There is no CPU fallback.
"""

from feinsum_b200.cl_utils import CudaDevice, CudaQueue, FakeCLDevice
from feinsum_b200.contraction_schedule import (
    get_opt_einsum_contraction_schedule,
    get_trivial_contraction_schedule,
)
from feinsum_b200.diagnostics import (
    CudaBackendError,
    InvalidParameterError,
    NoDevicePeaksInfoError,
    NoFactInDatabaseError,
    TransformValidationError,
)
from feinsum_b200.einsum import (
    Array,
    BatchedEinsum,
    EinsumAxisAccess,
    FreeAxis,
    SizeParam,
    SummationAxis,
)
from feinsum_b200.make_einsum import array, batched_einsum, einsum
from feinsum_b200.utils import IndexNameGenerator

__version__ = "2025.3+b200.r2"

# codegen / measure entry points (reference src/feinsum/__init__.py:37-68); imported lazily so
# that building the front-end objects does not need torch or the CUDA library
_LAZY = {
    "generate_cuda": "feinsum_b200.codegen.cuda",
    "CudaProgram": "feinsum_b200.codegen.cuda",
    "timeit": "feinsum_b200.measure",
    "measure_giga_op_rate": "feinsum_b200.measure",
    "get_roofline_flop_rate": "feinsum_b200.measure",
    "stringify_comparison_vs_roofline": "feinsum_b200.measure",
    "validate_batched_einsum_transform": "feinsum_b200.measure",
    "canonicalize_einsum": "feinsum_b200.canonicalization",
    "query": "feinsum_b200.sql_utils",
    "retrieve": "feinsum_b200.sql_utils",
    "record_facts": "feinsum_b200.sql_utils",
    "get_timed_einsums_in_db": "feinsum_b200.sql_utils",
    "DEFAULT_DB": "feinsum_b200.sql_utils",
    "autotune": "feinsum_b200.tuning",
}


# the reference's loopy-facing entry points (reference src/feinsum/__init__.py:3-6,19-24): there is
# no loopy in this backend -- the name resolves, the call says what to use instead
_LOOPY_ONLY = {
    "generate_loopy": "feinsum_b200.generate_cuda(einsum) returns the CudaProgram a transform acts on",
    "generate_loopy_with_opt_einsum_schedule":
        "feinsum_b200.generate_cuda(einsum, schedule=get_opt_einsum_contraction_schedule(einsum))",
    "get_a_matched_einsum": "build the BatchedEinsum with feinsum_b200.einsum / batched_einsum",
    "get_call_ids": "a CudaProgram holds one kernel; there are no loopy call ids",
    "identify_as_einsum": "build the BatchedEinsum with feinsum_b200.einsum / batched_einsum",
    "match_t_unit_to_einsum": "feinsum_b200.codegen.cuda.match_subscripts(einsum, pattern)",
}


def _loopy_only(name: str, instead: str):
    def stub(*args, **kwargs):
        raise NotImplementedError(
            f"feinsum_b200.{name}: the B200 backend has no loopy translation units "
            f"(SURVEY.md Appendix A); use {instead}")

    stub.__name__ = name
    stub.__doc__ = f"Not available in the CUDA backend (loopy-only in the reference); use {instead}."
    return stub


def __getattr__(name: str):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(_LAZY[name]), name)
    if name in _LOOPY_ONLY:
        return _loopy_only(name, _LOOPY_ONLY[name])
    raise AttributeError(f"module 'feinsum_b200' has no attribute '{name}'")
