#!/usr/bin/env python
"""Diffs the reference's preview PNGs (run_reference.sh) against tools/ghc_crosscheck/expected/.

Bar: every channel of every pixel within 1 LSB, and the number of pixels that differ at all is printed and
must stay below 0.5 % per image (GHC's `**`, `exp`, `cos` and this repository's libm may round differently in
the last place, which can move a value across a rounding boundary of toWord8).  Anything larger falsifies one
of the four third-party behaviours the oracle had to recall (oracle/oracle_thirdparty.c): massiv-io's HSI->RGB,
its toWord8 rounding, kdt's inRadius, linear's lookAt/normalize -- the report says which scenes, so the culprit
can be narrowed (no disk + black sky = camera only; stars = HSI + inRadius; disk = HSI of diskColor).

usage: python tools/ghc_crosscheck/compare.py REFERENCE_OUT_DIR"""
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref_dir = sys.argv[1]
    manifest = json.load(open(os.path.join(HERE, "expected", "manifest.json")))
    ok = True
    for name, meta in sorted(manifest["images"].items()):
        p = os.path.join(ref_dir, name)
        if not os.path.exists(p):
            print(f"{name}: MISSING in {ref_dir}")
            ok = False
            continue
        ref = np.asarray(Image.open(p).convert("RGB")).astype(int)
        exp = np.asarray(Image.open(os.path.join(HERE, "expected", name)).convert("RGB")).astype(int)
        if ref.shape != exp.shape:
            print(f"{name}: size {ref.shape[1]}x{ref.shape[0]} vs expected {exp.shape[1]}x{exp.shape[0]}  <-- prepareScene differs")
            ok = False
            continue
        d = np.abs(ref - exp)
        n_diff = int((d.max(axis=2) > 0).sum())
        frac = n_diff / (ref.shape[0] * ref.shape[1])
        verdict = "ok" if d.max() <= 1 and frac < 0.005 else "MISMATCH"
        ok = ok and verdict == "ok"
        where = ""
        if verdict != "ok":
            y, x = np.unravel_index(np.argmax(d.max(axis=2)), d.shape[:2])
            where = f"; worst pixel (x={x}, y={y}): reference {ref[y, x].tolist()} vs expected {exp[y, x].tolist()}"
        print(f"{name}: max |diff| {int(d.max())} LSB, {n_diff} pixels differ ({100 * frac:.3f} %) -> {verdict}{where}")
    print("PARITY PINNED: the oracle's recalled third-party behaviours reproduce the reference" if ok else
          "PARITY FALSIFIED: see the mismatches above and oracle/oracle_thirdparty.c")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
