"""N > 1 host logic on CPU: element ranges, local slicing and the optional result gather,
run with world_size 2 over gloo (no GPU).  The per-rank compute here is the numpy oracle --
test infrastructure standing in for the kernels, which need a GPU."""

import os
import socket

import numpy as np
import pytest

from feinsum_b200 import sharding
from tests import einsums as E


def test_element_ranges_partition_the_axis():
    for n in (0, 1, 15, 16, 17, 1000, 4_000_000, 31_999_999):
        for world in (1, 2, 4, 8):
            spans = [sharding.element_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
                assert hi == lo2 and lo <= hi
            for lo, hi in spans:
                if hi > lo:      # non-empty ranges start on a chunk boundary
                    assert lo % sharding.ALIGN == 0 and (hi % sharding.ALIGN == 0 or hi == n)
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) < 2 * sharding.ALIGN


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, q) -> None:
    import torch
    import torch.distributed as dist

    from oracle import np_oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for einsum in (E.grad(), E.div(), E.lift_fe(b=2)):
            glob = np_oracle.generate_input_arrays(einsum, n, 0)          # same on every rank
            loc = sharding.local_inputs(einsum, glob, world, rank)
            lo, hi = sharding.element_range(n, world, rank)
            loc_out = np_oracle.reference_outputs(einsum, loc)            # stands in for the kernel
            loc_t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in loc_out.items()}
            full = sharding.gather_outputs(einsum, loc_t, n)
            ref = np_oracle.reference_outputs(einsum, glob)
            for k in ref:
                np.testing.assert_allclose(full[k].numpy(), ref[k], rtol=1e-13)
            # max-over-ranks reduction used by bench.py
            t = torch.tensor([float(rank + 1)], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            assert t.item() == float(world)
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 1000])
def test_two_rank_shard_and_gather_over_gloo(n):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
