#!/usr/bin/env python
"""HBM bandwidth by read/write mix (torch kernels, CUDA events): write-only (fill_), copy (1:1), read-only (sum),
and a 1:3 read:write mix (one source broadcast to three destinations, the fused hex derivative's mix)."""
import json
import torch

dev = torch.device("cuda", 0)
n = 1 << 30                       # 8 GiB of float64
a = torch.rand(n, dtype=torch.float64, device=dev)
b = torch.empty_like(a)
c = torch.empty_like(a)
d = torch.empty_like(a)


def timeit(fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    e.synchronize()
    return nbytes * reps / (s.elapsed_time(e) * 1e-3) * 1e-9


out = {
    "write_only_fill_gbs": timeit(lambda: b.fill_(1.0), 8 * n),
    "copy_1r1w_gbs": timeit(lambda: b.copy_(a), 16 * n),
    "read_only_sum_gbs": timeit(lambda: a.sum(), 8 * n),
}


def one_to_three():
    torch.mul(a, 2.0, out=b)
    torch.mul(a, 3.0, out=c)
    torch.mul(a, 4.0, out=d)


out["three_copies_3r3w_gbs"] = timeit(one_to_three, 48 * n)
print(json.dumps(out))
