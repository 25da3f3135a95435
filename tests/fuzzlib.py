"""Random batched einsums and random isomorphic re-spellings (test helper).
Same distribution idea as the reference's fuzz (test/testlib.py:275-415): up to
16 rows, 8 operands, 14 indices, repeated indices inside operands, scalar
operands, per-(row, position) dtypes, operands re-used across rows."""

from __future__ import annotations

import numpy as np

import feinsum_b200 as f


def random_batched_einsum(rng: np.random.Generator, max_rank: int = 7):
    b = int(rng.integers(1, 17))
    n = int(rng.integers(1, 9))
    n_free = int(rng.integers(1, 8))
    n_redn = int(rng.integers(0, 8))
    letters = [chr(97 + (k + 8) % 26) for k in range(n_free + n_redn)]
    out_idx = letters[:n_free]
    while True:
        in_idx = [
            [letters[int(rng.integers(0, len(letters)))] for _ in range(int(rng.integers(0, max_rank + 1)))]
            for _ in range(n)
        ]
        if set(out_idx) <= {i for s in in_idx for i in s}:
            break
    length = {idx: int(rng.choice([4, 8, 16, 32, 64])) for idx in letters}
    dtypes = [[np.dtype(rng.choice(["float16", "float32", "float64"])) for _ in range(n)] for _ in range(b)]
    pool: dict[tuple, list[str]] = {}
    counter = 0
    rows = []
    for i in range(b):
        row = []
        for j, idxs in enumerate(in_idx):
            shape = tuple(length[k] for k in idxs)
            key = (shape, dtypes[i][j])
            names = pool.setdefault(key, [])
            if names and rng.random() < 0.7:
                name = names[int(rng.integers(0, len(names)))]
            else:
                name = f"A{counter // 26}{chr(65 + counter % 26)}"
                counter += 1
                names.append(name)
            row.append(f.array(name, shape, dtypes[i][j]))
        rows.append(row)
    sub = ",".join("".join(s) for s in in_idx) + "->" + "".join(out_idx)
    return f.batched_einsum(sub, rows)


def shuffled_copy(e, rng: np.random.Generator):
    names = sorted(e.all_args)
    idxs = sorted(e.all_indices)
    s_row = rng.permutation(e.b)
    s_pos = rng.permutation(e.n)
    s_idx = dict(zip(idxs, rng.permutation(idxs)))
    s_arg = dict(zip(names, rng.permutation(names)))
    sub = ",".join("".join(s_idx[i] for i in e.in_idx_sets[j]) for j in s_pos)
    sub += "->" + "".join(s_idx[i] for i in e.out_idx_set)
    rows = [[e.args[i][j].copy(name=str(s_arg[e.args[i][j].name])) for j in s_pos] for i in s_row]
    return f.batched_einsum(sub, rows)
