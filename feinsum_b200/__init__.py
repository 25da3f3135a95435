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

__version__ = "2025.3+b200.r1"

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


def __getattr__(name: str):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(_LAZY[name]), name)
    raise AttributeError(f"module 'feinsum_b200' has no attribute '{name}'")
