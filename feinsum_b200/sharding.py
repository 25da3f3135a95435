"""
Element-axis sharding over the GPUs of one box.

Every output entry of a DG einsum depends on exactly one element, so the
symbolic axis is split into contiguous ranges, one per rank, the operator
matrices are replicated, and there is **no collective on the data path**.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used for two
things only: the max-over-ranks time of a benchmark step and the *optional*
gather of results onto every rank.

Arrays whose element axis is not the leading one (``J(3,3,E)``, ``u(3,E,35)``,
``out(3,E,35)``) are not contiguous under a split of a global array; ranks
therefore own complete local arrays of extent ``E_local`` and a gather
concatenates along the element axis, slab by slab.
"""

from __future__ import annotations

from typing import Any

from feinsum_b200.einsum import BatchedEinsum, SizeParam

#: ranges start on multiples of this (the kernels' chunk) so every rank but the last runs whole chunks
ALIGN = 16


def element_range(n_elements: int, world_size: int, rank: int, align: int = ALIGN) -> tuple[int, int]:
    """Contiguous ``[lo, hi)`` owned by *rank*: sizes differ by at most one aligned block."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    n_blocks = -(-n_elements // align)
    base, extra = divmod(n_blocks, world_size)
    lo_b = rank * base + min(rank, extra)
    hi_b = lo_b + base + (1 if rank < extra else 0)
    return min(lo_b * align, n_elements), min(hi_b * align, n_elements)


def long_axis(shape: tuple[Any, ...]) -> int | None:
    """Position of the symbolic (element) axis in an operand / output shape, if any."""
    axes = [k for k, d in enumerate(shape) if isinstance(d, SizeParam)]
    if len(axes) > 1:
        raise NotImplementedError("more than one symbolic axis")
    return axes[0] if axes else None


def local_inputs(einsum: BatchedEinsum, arrays: dict[str, Any], world_size: int, rank: int) -> dict[str, Any]:
    """Slice every operand of a *global* problem down to this rank's element range
    (operands without the element axis are passed through)."""
    out: dict[str, Any] = {}
    n = None
    for name, shape in einsum.arg_to_shape.items():
        ax = long_axis(shape)
        if ax is not None:
            n = int(arrays[name].shape[ax])
    for name, shape in einsum.arg_to_shape.items():
        ax = long_axis(shape)
        a = arrays[name]
        if ax is None:
            out[name] = a
        else:
            lo, hi = element_range(n, world_size, rank)
            sl = [slice(None)] * a.ndim
            sl[ax] = slice(lo, hi)
            piece = a[tuple(sl)]
            out[name] = piece.contiguous() if hasattr(piece, "contiguous") else piece.copy()
    return out


def gather_outputs(einsum: BatchedEinsum, local_outs: dict[str, Any], n_elements: int,
                   group: Any = None) -> dict[str, Any]:
    """Optional result gather: every rank receives the full outputs (``all_gather`` of the
    uneven element ranges along the element axis).  Works with NCCL and gloo tensors."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    ax = long_axis(einsum.shape)
    full: dict[str, Any] = {}
    for name, loc in local_outs.items():
        if ax is None:
            full[name] = loc
            continue
        spans = [element_range(n_elements, world, r) for r in range(world)]
        width = max(hi - lo for lo, hi in spans)
        # all_gather wants equal shapes: pad the element axis to the widest range, trim after
        shp = list(loc.shape)
        shp[ax] = width
        padded = torch.zeros(shp, dtype=loc.dtype, device=loc.device)
        padded.narrow(ax, 0, loc.shape[ax]).copy_(loc)
        pieces = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(pieces, padded, group=group)
        full[name] = torch.cat([p.narrow(ax, 0, hi - lo) for p, (lo, hi) in zip(pieces, spans)], dim=ax)
    return full
