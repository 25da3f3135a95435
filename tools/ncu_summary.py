#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the handful of counters DESIGN.md cites.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source N]"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"), d.get("Block Size"), d.get("Grid Size"))
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
        stalls = {k: float(d[k]) for k in hdr if k.startswith("smsp__pcsamp_warps_issue_stalled_")
                  and not k.endswith("_not_issued") and d[k] not in ("", "n/a")}
        tot = sum(stalls.values()) or 1.0
        print("  stall samples:", ", ".join(f"{k[len('smsp__pcsamp_warps_issue_stalled_'):]}={v / tot:.1%}"
                                            for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.005))


def source(path, top):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    try:
        i_s = hdr.index("# Samples") if "# Samples" in hdr else next(i for i, h in enumerate(hdr) if "Samples" in h)
    except StopIteration:
        print(hdr); return
    i_src = hdr.index("Source")
    body = [r for r in rows[1:] if len(r) == len(hdr)]
    tot = sum(float(r[i_s] or 0) for r in body) or 1
    print("columns:", hdr[:12])
    ranked = sorted(enumerate(body), key=lambda ir: -float(ir[1][i_s] or 0))[:top]
    for i, r in sorted(ranked):
        print(f"  #{i:5d} {float(r[i_s] or 0) / tot:6.2%}  {r[i_src][:110]}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if "--source" in sys.argv:
        source(sys.argv[1], int(sys.argv[sys.argv.index("--source") + 1]))
