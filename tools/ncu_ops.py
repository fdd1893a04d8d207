#!/usr/bin/env python
"""Aggregates an ncu report's source page by SASS opcode: executed warp instructions, shared-memory
wavefronts and stall samples per kernel.  usage: python tools/ncu_ops.py report.ncu-rep [units]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed.sum",
            "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
    for r in rows[2:]:
        print("===", r[hdr.index("Kernel Name")][:70])
        for k in keys:
            if k in hdr:
                print(f"   {k} = {r[hdr.index(k)]} {rows[1][hdr.index(k)]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    hdr, agg, wf, smp, name = None, None, None, None, None

    def flush():
        if agg:
            tot = sum(agg.values())
            print(f"=== {name}: {tot} warp instructions = {tot / units:.1f} per unit; samples {sum(smp.values())}")
            for op, c in agg.most_common(28):
                print(f"   {op:24s} {c:>12d} {c / units:10.1f}/unit  shared wavefronts {wf[op] / units:9.1f}/unit  samples {smp[op]}")

    for r in csv.reader(io.StringIO(src)):
        if len(r) > 1 and r[0] == "Kernel Name":
            flush()
            name, agg, wf, smp, hdr = r[1][:70], collections.Counter(), collections.Counter(), collections.Counter(), None
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) // 2:
            continue
        d = dict(zip(hdr, r))
        toks = d["Source"].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.rstrip(";")
        agg[op] += int(d["Instructions Executed"] or 0)
        wf[op] += int(d.get("L1 Wavefronts Shared") or 0)
        smp[op] += int(d.get("# Samples") or 0)
    flush()


if __name__ == "__main__":
    main()
