#!/usr/bin/env python
"""Writes profiles/trace_kernel_traffic.json from an ncu --set full capture of the trace kernel on the headline
frame and the RK4 step count of that launch: DRAM bytes per launch and executed instructions per RK4 step
(FP64-pipe instructions = DFMA + DMUL + DADD + DSETP + ..., everything else).  bench.py reads the file.
usage: python tools/ncu_trace_summary.py REPORT.ncu-rep STEPS [OUT.json]"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = re.compile(r"^D(FMA|MUL|ADD|SETP|MNMX)\b")


def main():
    rep, steps = sys.argv[1], int(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "trace_kernel_traffic.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(k):
        v = float(r[hdr.index(k)])
        u = units[hdr.index(k)]
        return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    h, fp64, other = None, 0, 0
    ops = collections.Counter()
    for row in csv.reader(io.StringIO(src)):
        if row and row[0] == "Address":
            h = row
            continue
        if h is None or len(row) < len(h) // 2:
            continue
        d = dict(zip(h, row))
        toks = d["Source"].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).rstrip(";")
        n = int(d["Instructions Executed"] or 0)
        ops[op.split(".")[0]] += n
        if FP64.match(op):
            fp64 += n
        else:
            other += n
    warp_steps = steps / 32.0
    res = {
        "kernel": r[hdr.index("Kernel Name")],
        "capture": f"profiles/{os.path.basename(rep)} (ncu --set full --clock-control none), 4096x4096 default-aa frame, {steps} RK4 steps",
        "gpu_time_ms": val("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)
        if units[hdr.index("gpu__time_duration.sum")] in ("ms", "us", "ns", "s") else None,
        "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
        "algorithmic_bytes": 268435456,
        "fp64_instr_per_rk4_step": fp64 / warp_steps, "other_instr_per_rk4_step": other / warp_steps,
        "fp64_pipe_active_pct": float(r[hdr.index("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")]),
        "issue_active_pct": float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
        "registers_per_thread": float(r[hdr.index("launch__registers_per_thread")]),
        "top_opcodes_per_rk4_step": {k: v / warp_steps for k, v in ops.most_common(12)},
    }
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
