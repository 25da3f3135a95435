"""The C-ABI library builds, loads and exports everything include/fnsm_b200.h declares."""

import os
import re

import pytest

from feinsum_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        from feinsum_b200._build import build

        build()
    return _cabi.lib()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fnsm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fnsm_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(names) == set(_cabi.EXPORTED_SYMBOLS)


def test_abi_version_and_strerror(lib):
    assert lib.fnsm_b200_abi_version() == 1
    assert _cabi.strerror(0) == "success"
    assert "configuration" in _cabi.strerror(-3)
    assert "unknown" in _cabi.strerror(-99)
    assert "invalid" in _cabi.strerror(1).lower()  # cudaErrorInvalidValue


def test_cfg_space_introspection(lib):
    space = _cabi.query_cfg_space(_cabi.K_DIV)
    names = [p["name"] for p in space]
    assert "variant" in names and "ctas_per_sm" in names
    for p in space:
        assert p["lo"] <= p["hi"] and p["step"] >= 1
    assert _cabi.query_cfg_space(_cabi.K_GENERIC) == []
    assert [p["name"] for p in _cabi.query_cfg_space(_cabi.K_TENSOR_PRODUCT)] == ["ctas_per_sm"]


def test_struct_layouts_match_header(lib):
    import ctypes as C

    assert C.sizeof(_cabi.Cfg) == 32
    assert C.sizeof(_cabi.CfgRange) == 40
    assert C.sizeof(_cabi.EinsumDesc) == 16 + 8 * 12 * (2 + 6)
    assert C.sizeof(_cabi.WaveArgs) == 8 * (6 + 4 + 2 + 4)


def test_bad_config_maps_to_invalid_parameter():
    import feinsum_b200 as f

    with pytest.raises(f.InvalidParameterError):
        _cabi.check(-3, "x")
    with pytest.raises(f.CudaBackendError):
        _cabi.check(-1, "x")
    with pytest.raises(f.InvalidParameterError):
        _cabi.make_cfg({"no_such_knob": 1})


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import feinsum_b200 as f

    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    _cabi.lib.cache_clear()
    try:
        with pytest.raises(f.CudaBackendError):
            _cabi.lib()
    finally:
        monkeypatch.undo()
        _cabi.lib.cache_clear()


def test_se_shape_table_mirrors_the_library(lib):
    import numpy as np

    from feinsum_b200.codegen.cuda import se_kernel_available

    for dt, code in ((np.dtype("float64"), _cabi.FNSM_F64), (np.dtype("float32"), _cabi.FNSM_F32)):
        for ns in range(1, 6):
            for ni in (3, 4, 6, 10, 15, 20, 35, 36):
                for nj in (ni, ni + 1):
                    assert bool(lib.fnsm_b200_opmat_se_supported(code, ns, ni, nj)) == \
                        se_kernel_available(dt, ns, ni, nj), (dt, ns, ni, nj)


def test_make_cfg_measurement_switches():
    """``with_params`` keys that land in ``fnsm_cfg.reserved`` (include/fnsm_b200.h): flags -> [0], stagger -> [1],
    dbgk -> bits 0-3 of [2], fast_start -> bits 4-5 of [2]; they do not disturb each other."""
    from feinsum_b200 import _cabi
    from feinsum_b200.diagnostics import InvalidParameterError

    assert _cabi.make_cfg(None) is None and _cabi.make_cfg({}) is None
    cfg = _cabi.make_cfg({"threads": 384, "variant": 1, "flags": 1, "stagger": 3000, "dbgk": 5, "fast_start": 2}).contents
    assert (cfg.threads, cfg.variant) == (384, 1)
    assert list(cfg.reserved) == [1, 3000, 5 | (2 << 4)]
    cfg = _cabi.make_cfg({"fast_start": 1, "dbgk": 0}).contents
    assert list(cfg.reserved) == [0, 0, 1 << 4]
    cfg = _cabi.make_cfg({"dbgk": 7, "fast_start": 0}).contents
    assert list(cfg.reserved) == [0, 0, 7]
    with pytest.raises(InvalidParameterError):
        _cabi.make_cfg({"no_such_knob": 1})
