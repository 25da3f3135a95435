#!/usr/bin/env python
"""Populate the facts database shipped with the package (``feinsum_b200/data/cuda_facts_v1.sqlite``) by
running the autotuner over the BASELINE einsums on the current GPU -- the CUDA counterpart of the
reference's ``data/transform_archive_v5.sqlite`` (reference ``tuning/__init__.py:485-633`` fills that one).

    python tools/populate_db.py [--db PATH] [--elements 4000000] [--secs 0.15]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import feinsum_b200 as f  # noqa: E402
from feinsum_b200 import measure, sql_utils, tuning  # noqa: E402
from tests import einsums as E  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--db", default=sql_utils.DEFAULT_DB)
    ap.add_argument("--elements", type=int, nargs="+", default=[4_000_000, 100_000])
    ap.add_argument("--secs", type=float, default=0.15)
    ap.add_argument("--only", default="", help="restrict to launch-space modules whose file name contains this")
    ap.add_argument("--dtypes", nargs="+", default=["float64", "float32"])
    ap.add_argument("--replace", action="store_true",
                    help="delete the selected einsums' facts at the selected sizes first (re-measure after a kernel change)")
    args = ap.parse_args()
    measure.N_MIN_SIM_SECS = args.secs
    cq = f.CudaQueue(0)
    impls = os.path.join(os.path.dirname(os.path.abspath(tuning.__file__)), "impls")
    cases = []
    for dt in args.dtypes:
        cases += [(E.grad(dtype=dt), "xre_rij_ej_to_xei.py"), (E.div(dtype=dt), "xre_rij_xej_to_ei.py"),
                  (E.lift_fe(dtype=dt), "ifj_fe_fej_to_ei.py"), (E.lift_ef(dtype=dt), "ef_fij_fej_to_ei.py"),
                  (E.tensor_product(0, 8, dt), "eabc_ia_to_eibc.py")]
    cases = [(e, mod) for e, mod in cases if args.only in mod]
    if args.replace:
        import sqlite3

        from feinsum_b200.canonicalization import canonicalize_einsum

        con = sqlite3.connect(args.db)
        for n in args.elements:
            for e, _ in cases:
                key = sql_utils._key(canonicalize_einsum(e), cq.device)
                con.execute(f"DELETE FROM {sql_utils.TIMINGS_TABLENAME} WHERE (subscripts = ? AND index_to_length = ? AND "
                            "args = ? AND arg_to_dtype = ? AND device_name = ? AND n_elements = ?);", (*key, n))
        con.commit()
        con.close()
    for n in args.elements:
        for e, mod in cases:
            best = tuning.autotune(e, os.path.join(impls, mod), cq, db_path=args.db, long_dim_length=n)
            q = f.query(e, cq.device, database=args.db)
            t = min(x.runtime_in_sec for x in q if x.n_elements == n)
            print(f"{mod:24s} {next(iter(e.arg_to_dtype.values()))!s:8s} E={n:8d}  best={best}  {t * 1e3:.4f} ms", flush=True)


if __name__ == "__main__":
    main()
