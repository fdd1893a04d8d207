#!/usr/bin/env python
"""Writes what this repository expects the REAL reference to produce, for someone who has GHC:

  tools/ghc_crosscheck/expected/prev-<scene>.png   the nine shipped scenes as `blackstar --preview` renders
                                                   them (app/Main.hs:93-103: long side 300, no supersampling,
                                                   no bloom), 8-bit sRGB PNG, from the synthetic catalogue
  tools/ghc_crosscheck/expected/manifest.json      per image: size, sha256 of the RGB8 bytes, RK4 step count

The images come from oracle/ (the C restatement of the reference, CPU) -- no GPU needed to regenerate them --
and tests/test_gpu_parity.py::test_ghc_crosscheck_expectations checks that the CUDA path renders the same
bytes (<= 1 LSB on a counted handful of pixels).  run_reference.sh produces the other side on a machine
with `stack`; compare.py diffs the two.
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from blackstar_b200 import config, starmap  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

SCENES = ["closeup", "default", "default-aa", "fartheraway", "lensing-disk", "lensing", "wideangle-disk", "wideangle", "wideangle1"]


def main():
    out = os.path.join(HERE, "expected")
    os.makedirs(out, exist_ok=True)
    stars = starmap.synthetic_stars()          # N = 468 861, seed 20190412: the bytes make_inputs.py writes as stars.ppm
    tree = po.Tree(stars)
    manifest = {"catalogue": {"n": int(len(stars)), "seed": starmap.DEFAULT_SEED,
                              "sha256_ppm": hashlib.sha256(starmap.synthetic_catalogue()).hexdigest()}, "images": {}}
    for name in SCENES:
        cfg = config.prepare_scene(config.load_config(os.path.join(ROOT, "scenes", name + ".yaml")), True)
        img, steps = po.render(cfg, tree)
        if cfg.scene.bloomStrength != 0:
            img = po.bloom(cfg.scene.bloomStrength, cfg.scene.bloomDivider, img)
        rgb8 = po.to_srgb8(img)
        Image.fromarray(rgb8).save(os.path.join(out, f"prev-{name}.png"), format="PNG", optimize=True)
        manifest["images"][f"prev-{name}.png"] = {"width": int(rgb8.shape[1]), "height": int(rgb8.shape[0]),
                                                  "sha256_rgb8": hashlib.sha256(rgb8.tobytes()).hexdigest(), "rk4_steps": int(steps)}
        print(name, rgb8.shape, steps)
    json.dump(manifest, open(os.path.join(out, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
