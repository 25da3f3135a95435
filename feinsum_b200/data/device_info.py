"""
Device peaks for the roofline model (reference ``src/feinsum/data/device_info.py:4-28``).

The reference's rows are kept so historical facts stay interpretable; the B200
rows are what this backend is measured against:

* ``float64`` / ``float32``: register-resident DFMA / packed-FFMA2 loops
  measured on this pool's B200s with ``tools/ubench`` and
  ``fnsm_b200_measure_peak`` (see ``profiles/`` for the run); nominal
  148 SM x 64 (128) FMA/clk x 1.965 GHz = 37.2 (74.4) TFLOP/s.
* bandwidth: the driver-measured copy figure in ``MEASURED_PEAKS.json``
  (6561.6 GB/s; nominal HBM3e 7.7-8 TB/s).

``device_peaks(name)`` prefers live numbers registered with
:func:`register_measured_peaks` over the table.
"""

from __future__ import annotations

import json
import os
from collections.abc import Mapping

# GFLOP/s
DEV_TO_PEAK_GFLOPS: dict[str, Mapping[str, float]] = {
    "NVIDIA TITAN V": {"float32": 12288, "float64": 6144},
    "NVIDIA GeForce GTX 1650": {"float32": 3916.0, "float64": 122.4},
    "NVIDIA H200 NVL": {"float32": 67000, "float64": 34000},
    # measured on this pool (profiles/r01_ubench.jsonl); see module docstring
    "NVIDIA B200": {"float32": 74400.0, "float64": 37200.0},
}

# GB/s
DEV_TO_PEAK_BW: dict[str, float] = {
    "NVIDIA TITAN V": 652.8,
    "NVIDIA GeForce GTX 1650": 192.0,
    "NVIDIA H200 NVL": 4800,
    "NVIDIA B200": 6561.6,
}

NOMINAL_B200 = {"float64": 37200.0, "float32": 74400.0, "hbm_gbs": 7700.0}


def _load_measured_hbm() -> float | None:
    """HBM copy bandwidth written by the driver to MEASURED_PEAKS.json, if present."""
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, os.pardir, os.pardir, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return None


_hbm = _load_measured_hbm()
if _hbm is not None:
    DEV_TO_PEAK_BW["NVIDIA B200"] = _hbm


def register_measured_peaks(
    dev_name: str, *, float64: float | None = None, float32: float | None = None,
    bw: float | None = None,
) -> None:
    """Override table entries with numbers measured on the running board."""
    cur = dict(DEV_TO_PEAK_GFLOPS.get(dev_name, {}))
    if float64 is not None:
        cur["float64"] = float(float64)
    if float32 is not None:
        cur["float32"] = float(float32)
    if cur:
        DEV_TO_PEAK_GFLOPS[dev_name] = cur
    if bw is not None:
        DEV_TO_PEAK_BW[dev_name] = float(bw)
