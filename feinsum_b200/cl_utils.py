"""
Device / queue handles.

The reference passes a ``pyopencl.CommandQueue`` (``cq``) to everything that
launches a kernel and "anything with a ``.name``" to everything that only
looks things up (reference ``src/feinsum/cl_utils.py:9-21``,
``measure.py:197-204``, ``sql_utils.py:160-176``).  The file keeps its
reference name so imports line up; the queue is now a CUDA stream on one
B200:

* :class:`DeviceT`       -- protocol: has ``.name``
* :class:`FakeCLDevice`  -- name-only stand-in for database look-ups
* :class:`CudaDevice`    -- a real GPU (``name``, ``vendor``, ``driver_version``)
* :class:`CudaQueue`     -- ``cq`` replacement: ``.device``, ``.stream``
  (raw ``cudaStream_t`` as int), ``.finish()``

torch is used here only to own the CUDA context, streams and memory.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Protocol


class DeviceT(Protocol):
    """Anything that looks like ``pyopencl.Device`` as far as feinsum cares."""

    @property
    def name(self) -> str: ...


@dataclass(frozen=True, repr=True, eq=True)
class FakeCLDevice:
    name: str


@dataclass(frozen=True, repr=True, eq=True)
class CudaDevice:
    index: int
    name: str
    vendor: str = "NVIDIA"
    driver_version: str = ""
    sm_count: int = 0
    cc: tuple[int, int] = (0, 0)

    @staticmethod
    def from_index(index: int = 0) -> "CudaDevice":
        import torch

        if not torch.cuda.is_available():
            from feinsum_b200.diagnostics import CudaBackendError

            raise CudaBackendError(
                "No CUDA device is visible; feinsum_b200 has no CPU fallback."
            )
        props = torch.cuda.get_device_properties(index)
        try:
            import pynvml

            pynvml.nvmlInit()
            drv = pynvml.nvmlSystemGetDriverVersion()
            drv = drv.decode() if isinstance(drv, bytes) else str(drv)
        except Exception:  # noqa: BLE001
            drv = "unknown"
        return CudaDevice(
            index=index,
            name=props.name,
            driver_version=f"driver{drv}-cuda{torch.version.cuda}",
            sm_count=props.multi_processor_count,
            cc=(props.major, props.minor),
        )


class CudaQueue:
    """In-order execution queue = one CUDA stream on one device.

    ``CudaQueue()`` wraps torch's *current* stream of ``cuda:index`` so that
    ``torch.cuda.Event`` timing and the native launches see the same stream.
    """

    def __init__(self, device: int | CudaDevice = 0, stream: Any | None = None):
        import torch

        self.device = (
            device if isinstance(device, CudaDevice) else CudaDevice.from_index(device)
        )
        self.torch_device = torch.device("cuda", self.device.index)
        self._torch_stream = (
            stream if stream is not None else torch.cuda.current_stream(self.torch_device)
        )

    @property
    def torch_stream(self) -> Any:
        return self._torch_stream

    @property
    def stream(self) -> int:
        """Raw ``cudaStream_t`` handle (0 = legacy default stream)."""
        return int(self._torch_stream.cuda_stream)

    def finish(self) -> None:
        self._torch_stream.synchronize()

    def __repr__(self) -> str:
        return f"CudaQueue({self.device.name!r}, stream=0x{self.stream:x})"


def as_queue(cq: Any) -> CudaQueue:
    """Accept a :class:`CudaQueue`, a device index, or ``None`` (``cuda:0``)."""
    if isinstance(cq, CudaQueue):
        return cq
    if cq is None:
        return CudaQueue(0)
    if isinstance(cq, int):
        return CudaQueue(cq)
    raise TypeError(f"expected a CudaQueue or device index, got {type(cq).__name__}")
