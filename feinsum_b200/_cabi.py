"""
ctypes binding of ``include/fnsm_b200.h`` -- the only door into the kernels.

Loading fails loudly (:class:`~feinsum_b200.diagnostics.CudaBackendError`)
when ``libfnsm_b200.so`` is absent; there is no fallback implementation.
"""

from __future__ import annotations

import ctypes as C
import os
from functools import cache
from typing import Any

from feinsum_b200.diagnostics import CudaBackendError, InvalidParameterError

#: FNSM_B200_LIB points the loader at another build of the library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("FNSM_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                           "libfnsm_b200.so")

FNSM_F64, FNSM_F32, FNSM_I32, FNSM_I64, FNSM_C64, FNSM_C128 = range(6)
OP_GRAD, OP_DIV, OP_LIFT_EF, OP_LIFT_FE = 0, 1, 2, 3
K_GENERIC, K_GRAD, K_DIV, K_LIFT, K_WAVE3D, K_TENSOR_PRODUCT, K_SE, K_HEX_DERIV = range(8)
MAX_INDICES, MAX_OPERANDS = 12, 6
E_BAD_CONFIG = -3

EXPORTED_SYMBOLS = (
    "fnsm_b200_generic_einsum",
    "fnsm_b200_opmat_batch",
    "fnsm_b200_opmat_se",
    "fnsm_b200_opmat_se_supported",
    "fnsm_b200_wave3d_fused",
    "fnsm_b200_tensor_product",
    "fnsm_b200_hex_deriv",
    "fnsm_b200_query_cfg_space",
    "fnsm_b200_measure_peak",
    "fnsm_b200_copy2d_async",
    "fnsm_b200_launch_count",
    "fnsm_b200_abi_version",
    "fnsm_b200_strerror",
)


class Cfg(C.Structure):
    _fields_ = [
        ("variant", C.c_int32),
        ("tile_e", C.c_int32),
        ("threads", C.c_int32),
        ("stages", C.c_int32),
        ("ctas_per_sm", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class CfgRange(C.Structure):
    _fields_ = [
        ("name", C.c_char * 24),
        ("lo", C.c_int32),
        ("hi", C.c_int32),
        ("step", C.c_int32),
        ("dflt", C.c_int32),
    ]


class EinsumDesc(C.Structure):
    _fields_ = [
        ("n_free", C.c_int32),
        ("n_sum", C.c_int32),
        ("n_operands", C.c_int32),
        ("dtype", C.c_int32),
        ("extent", C.c_int64 * MAX_INDICES),
        ("out_stride", C.c_int64 * MAX_INDICES),
        ("in_stride", (C.c_int64 * MAX_INDICES) * MAX_OPERANDS),
    ]


class WaveArgs(C.Structure):
    _fields_ = [
        ("J", C.c_void_p),
        ("D", C.c_void_p),
        ("v", C.c_void_p),
        ("u", C.c_void_p),
        ("L", C.c_void_p),
        ("Jface", C.c_void_p),
        ("F", C.c_void_p * 4),
        ("div_out", C.c_void_p),
        ("grad_out", C.c_void_p),
        ("lift_out", C.c_void_p * 4),
    ]


@cache
def lib() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise CudaBackendError(
            f"{LIB_PATH} is missing: build it with `python -m feinsum_b200._build` "
            "(needs nvcc). feinsum_b200 has no CPU fallback."
        )
    try:
        handle = C.CDLL(LIB_PATH)
    except OSError as exc:
        raise CudaBackendError(f"cannot load {LIB_PATH}: {exc}") from exc
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    pvp = C.POINTER(C.c_void_p)
    handle.fnsm_b200_generic_einsum.argtypes = [C.POINTER(EinsumDesc), i32, pvp, pvp, vp]
    handle.fnsm_b200_opmat_batch.argtypes = [
        i32, i32, vp, vp, pvp, pvp, i32, i32, i32, i32, i64, C.POINTER(Cfg), vp]
    handle.fnsm_b200_opmat_se.argtypes = [i32, i32, pvp, vp, pvp, pvp, i32, i32, i32, i32, i64, C.POINTER(Cfg), vp]
    handle.fnsm_b200_opmat_se_supported.argtypes = [i32, i32, i32, i32]
    handle.fnsm_b200_wave3d_fused.argtypes = [i32, C.POINTER(WaveArgs), i64, C.POINTER(Cfg), vp]
    handle.fnsm_b200_tensor_product.argtypes = [i32, vp, vp, vp, i32, i32, i64, C.POINTER(Cfg), vp]
    handle.fnsm_b200_hex_deriv.argtypes = [i32, vp, pvp, pvp, i32, i64, C.POINTER(Cfg), vp]
    handle.fnsm_b200_query_cfg_space.argtypes = [i32, C.POINTER(CfgRange), i32]
    handle.fnsm_b200_measure_peak.argtypes = [i32, C.POINTER(C.c_double)]
    handle.fnsm_b200_copy2d_async.argtypes = [vp, i64, vp, i64, i64, i64, i32, vp]
    handle.fnsm_b200_launch_count.argtypes = []
    handle.fnsm_b200_launch_count.restype = i64
    handle.fnsm_b200_abi_version.argtypes = []
    handle.fnsm_b200_strerror.argtypes = [C.c_int]
    handle.fnsm_b200_strerror.restype = C.c_char_p
    for name in EXPORTED_SYMBOLS:
        fn = getattr(handle, name)
        if fn.restype is C.c_int and name not in ("fnsm_b200_launch_count",):
            fn.restype = C.c_int
    if handle.fnsm_b200_abi_version() != 1:
        raise CudaBackendError("libfnsm_b200.so: ABI version mismatch")
    return handle


def strerror(code: int) -> str:
    return lib().fnsm_b200_strerror(int(code)).decode()


def check(code: int, what: str = "") -> None:
    """Raise on a non-zero return code of an ABI call."""
    if code == 0:
        return
    msg = f"{what}: {strerror(code)} (code {code})" if what else f"{strerror(code)} (code {code})"
    if code == E_BAD_CONFIG:
        raise InvalidParameterError(msg)
    raise CudaBackendError(msg)


def make_cfg(params: dict[str, Any] | None) -> Any:
    """``{"tile_e": 64, ...}`` -> ``fnsm_cfg*`` (or NULL for defaults)."""
    if not params:
        return None
    cfg = Cfg()
    known = {"variant", "tile_e", "threads", "stages", "ctas_per_sm"}
    for key, val in params.items():
        if key == "flags":  # debug bits (bit 0: force the plain-load path of the dmma kernels)
            cfg.reserved[0] = int(val)
            continue
        if key == "dbgk":  # dmma div: compile-time profiling variants (results invalid)
            cfg.reserved[2] = (cfg.reserved[2] & ~15) | (int(val) & 15)
            continue
        if key == "fast_start":  # dmma kernels: 1 / 2 force the small-launch instantiations on / off (0: by size)
            cfg.reserved[2] = (cfg.reserved[2] & 15) | ((int(val) & 3) << 4)
            continue
        if key == "stagger":  # dmma kernels: start-up phase offset (cycles) between warps of one sub-partition
            cfg.reserved[1] = int(val)
            continue
        if key not in known:
            raise InvalidParameterError(f"unknown launch parameter '{key}'")
        setattr(cfg, key, int(val))
    return C.pointer(cfg)


def launch_count() -> int:
    return int(lib().fnsm_b200_launch_count())


def query_cfg_space(kernel_id: int) -> list[dict[str, Any]]:
    buf = (CfgRange * 16)()
    n = lib().fnsm_b200_query_cfg_space(kernel_id, buf, 16)
    if n < 0:
        check(n, "query_cfg_space")
    return [
        {"name": buf[i].name.decode(), "lo": buf[i].lo, "hi": buf[i].hi,
         "step": buf[i].step, "default": buf[i].dflt}
        for i in range(min(n, 16))
    ]


def measure_peak(which: int) -> float:
    out = C.c_double(0.0)
    check(lib().fnsm_b200_measure_peak(which, C.byref(out)), "measure_peak")
    return float(out.value)
