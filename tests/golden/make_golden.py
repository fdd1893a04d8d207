"""Generates tests/golden/*.npz from the ORACLE (oracle/liboracle.so).

The reference itself (Haskell) cannot be run in this image (no ghc/stack/cabal), so these
are oracle outputs, not reference outputs; they pin the oracle against drift and give the
GPU tests a fixture that does not need the oracle at run time.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from blackstar_b200 import config, starmap  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

GOLDEN_STARS = dict(n=30000, seed=20190412)
WIDTH = 48


def golden_config(scene_path):
    cfg = config.load_config(scene_path)
    w, h = cfg.scene.resolution
    return config.with_resolution(cfg, WIDTH, max(8, WIDTH * h // w))


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    stars = starmap.synthetic_stars(**GOLDEN_STARS)
    tree = po.Tree(stars)
    scenes = sorted(f for f in os.listdir(os.path.join(ROOT, "scenes")) if f.endswith(".yaml"))
    data = {}
    for f in scenes:
        cfg = golden_config(os.path.join(ROOT, "scenes", f))
        img, steps = po.render(cfg, tree)
        bl = po.bloom(cfg.scene.bloomStrength, cfg.scene.bloomDivider, img)
        name = f[:-5]
        data[name + "/render"] = img
        data[name + "/bloomed"] = bl
        data[name + "/steps"] = np.array([steps], dtype=np.int64)
        data[name + "/srgb8"] = po.to_srgb8(bl)
    np.savez_compressed(os.path.join(out_dir, "scenes_48.npz"), **data)
    print("wrote", os.path.join(out_dir, "scenes_48.npz"), len(data), "arrays")


if __name__ == "__main__":
    main()
