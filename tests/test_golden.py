"""The oracle must keep reproducing the committed golden fixtures (tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

from blackstar_b200 import starmap  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

SCENES = ["closeup", "default", "default-aa", "fartheraway", "lensing-disk", "lensing", "wideangle-disk",
          "wideangle", "wideangle1"]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "scenes_48.npz"))


@pytest.mark.parametrize("scene", SCENES)
def test_oracle_reproduces_golden(golden, scenes_dir, scene):
    cfg = make_golden.golden_config(f"{scenes_dir}/{scene}.yaml")
    stars = starmap.synthetic_stars(**make_golden.GOLDEN_STARS)
    img, steps = po.render(cfg, po.Tree(stars))
    # libm may differ by an ulp between machines (sin/cos/exp); everything else is bit-exact
    np.testing.assert_allclose(img, golden[scene + "/render"], rtol=0, atol=1e-13)
    assert steps == int(golden[scene + "/steps"][0])
    bl = po.bloom(cfg.scene.bloomStrength, cfg.scene.bloomDivider, img)
    np.testing.assert_allclose(bl, golden[scene + "/bloomed"], rtol=0, atol=1e-13)
