#!/usr/bin/env python
"""Per-SASS-instruction stall samples of one profiled kernel, aggregated into
instruction classes and the hottest instructions.  usage: ncu_source.py rep [topN]"""
import csv, io, re, subprocess, sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(float(r[iS] or 0) for r in body) or 1
cls = Counter(); cnt = Counter()
for r in body:
    op = r[iSrc].split()[0] if r[iSrc].split() else "?"
    if op.startswith("@"):
        op = r[iSrc].split()[1]
    op = op.split(".")[0]
    cls[op] += float(r[iS] or 0); cnt[op] += float(r[iEx] or 0)
print("samples by opcode (share of samples | warp-instructions executed):")
for op, v in cls.most_common(14):
    print(f"  {op:10s} {v / tot:6.1%}  {cnt[op]:.3g}")
print(f"hottest {top} instructions (index, share, source, dominant stall):")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ranked = sorted(range(len(body)), key=lambda i: -float(body[i][iS] or 0))[:top]
for i in sorted(ranked):
    r = body[i]
    st = max(stall_cols, key=lambda c: float(r[c] or 0))
    print(f"  #{i:5d} {float(r[iS] or 0) / tot:6.2%}  {r[iSrc].strip()[:70]:70s} {hdr[st]}")
