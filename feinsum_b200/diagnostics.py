"""
Exception types of the batched-einsum backend.

Same names and base classes as the reference (``src/feinsum/diagnostics.py:33-65``)
so ``except`` clauses written against feinsum keep working; plus
:class:`CudaBackendError` for failures of the native sm_100a library.
"""


class EinsumTunitMatchError(ValueError):
    """A kernel/pattern could not be matched against an einsum."""


class InvalidParameterError(ValueError):
    """A launch configuration lies in the declared space but is illegal
    (the tuner maps it to ``time = inf``; reference ``tuning/__init__.py:557-559``)."""


class NoDevicePeaksInfoError(LookupError):
    """No roofline peaks are tabulated for the queried device."""


class TransformValidationError(RuntimeError):
    """The kernel's results differ from the ``numpy.einsum`` oracle
    (reference ``measure.py:186-192``)."""


class NoFactInDatabaseError(RuntimeError):
    """:func:`~feinsum_b200.sql_utils.retrieve` found no recorded fact."""


class CudaBackendError(RuntimeError):
    """The native library is missing, failed to load, or a launch returned an
    error code.  There is deliberately no CPU fallback behind this."""
