#!/usr/bin/env python
"""Writes the inputs of the cross-check into a directory: stars.ppm (the synthetic catalogue in the exact
PPM binary layout StarMap.readMap parses, src/StarMap.hs:45-58) and a copy of the nine scene files.
usage: python tools/ghc_crosscheck/make_inputs.py OUTDIR"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from blackstar_b200 import starmap  # noqa: E402  (numpy only; no GPU, no CUDA library needed)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "crosscheck_inputs"
    os.makedirs(os.path.join(out, "scenes"), exist_ok=True)
    data = starmap.synthetic_catalogue()
    with open(os.path.join(out, "stars.ppm"), "wb") as f:
        f.write(data)
    for name in sorted(os.listdir(os.path.join(ROOT, "scenes"))):
        shutil.copy(os.path.join(ROOT, "scenes", name), os.path.join(out, "scenes", name))
    print(f"{out}/stars.ppm: {len(data)} bytes, sha256 {hashlib.sha256(data).hexdigest()}")
    print(f"{out}/scenes: {len(os.listdir(os.path.join(out, 'scenes')))} scene files")


if __name__ == "__main__":
    main()
