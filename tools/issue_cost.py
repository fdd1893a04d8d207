#!/usr/bin/env python
"""Static issue-cost of the trace kernel's RK4 loop, read from the SASS of libblackstar_b200.so.

On B200 an FP64 instruction holds a sub-partition's issue port for two cycles and any other
instruction for one; `2*FP64 + others` per RK4 step reproduces the measured cycles of the kernel to
3 % (profiles/README.md).  This script finds the innermost loop with 8 MUFU.RSQ64H (two unrolled
steps) in every trace_tiles_kernel instantiation and prints that cost -- no GPU needed, so a change
that bloats the loop is caught on the CPU box (tests/test_abi_and_host.py uses it).
"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "blackstar_b200", "libblackstar_b200.so")
FP64 = re.compile(r"^D(FMA|MUL|ADD|SETP|MNMX)\b")


def loops(lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for chunk in sass.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0].strip()
        if "trace_tiles_kernel" not in name:
            continue
        ins = []
        for line in chunk.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())))
        best = None
        for k, (a, text) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < a:
                start = next(i for i, (aa, _) in enumerate(ins) if aa == int(m.group(1), 16))
                body = [t for _, t in ins[start:k + 1]]
                if sum("MUFU.RSQ64H" in t for t in body) == 8 and (best is None or len(body) < len(best)):
                    best = body
        if best:
            ops = Counter(t.split()[0].split(".")[0] for t in best)
            fp64 = sum(1 for t in best if FP64.match(t))
            out[name] = {"instructions": len(best), "fp64": fp64, "others": len(best) - fp64,
                         "cycles_per_step": (2 * fp64 + len(best) - fp64) / 2.0, "ops": dict(ops)}
    return out


if __name__ == "__main__":
    res = loops(sys.argv[1] if len(sys.argv) > 1 else LIB)
    for name, r in sorted(res.items()):
        short = re.sub(r"^_ZN3bsb18", "", name)[:40]
        print(f"{short:40s} 2 steps: {r['instructions']:4d} instr = {r['fp64']} FP64 + {r['others']} others"
              f"  -> {r['cycles_per_step']:.1f} issue cycles per RK4 step")
