"""A SECOND restatement of the four third-party behaviours the oracle had to recall
(oracle/oracle_thirdparty.c), written independently in numpy from the published definitions, so that a
transcription slip in the C file (which the CUDA code shares its formulas with) cannot hide:

* HSI -> RGB: the sector formula of Gonzalez & Woods, "Digital Image Processing", eq. 6.2-5..7, which is what
  massiv-io's ``toPixelRGB`` for ``HSI`` documents itself as (hue normalised to [0, 1), I = mean of R, G, B);
* ``toWord8`` for ``Double``: clamp to [0, 1], scale by 255, Haskell ``round`` (banker's rounding);
* ``Linear.Projection.lookAt`` / ``Linear.Metric.normalize``: the definitions in linear's haddocks;
* ``Data.KdMap.Static.inRadius``: "all points within the given radius" (closed ball) -> brute force.

None of this can prove what the Hackage packages actually do (that needs GHC: tools/ghc_crosscheck), but
it breaks the common mode between the oracle and the product.  Also pins the committed cross-check
expectations to the oracle, so they cannot drift apart unnoticed."""
import hashlib
import json
import os

import numpy as np
import pytest

from blackstar_b200 import config, starmap
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hsi_to_rgb_textbook(hue01, s, i):
    """Gonzalez & Woods: H in degrees; RG sector 0 <= H < 120, GB sector 120 <= H < 240, BR sector."""
    H = 360.0 * hue01
    def chroma(Hs):  # I * (1 + S cos(H) / cos(60 - H)) with angles in degrees
        return i * (1.0 + s * np.cos(np.radians(Hs)) / np.cos(np.radians(60.0 - Hs)))
    if 0.0 <= H < 120.0:
        b = i * (1.0 - s); r = chroma(H); g = 3.0 * i - (r + b)
    elif 120.0 <= H < 240.0:
        r = i * (1.0 - s); g = chroma(H - 120.0); b = 3.0 * i - (r + g)
    elif 240.0 <= H < 360.0:
        g = i * (1.0 - s); b = chroma(H - 240.0); r = 3.0 * i - (g + b)
    else:
        raise ValueError("hue outside [0, 1)")
    return np.array([r, g, b])


def test_hsi_to_rgb_against_the_textbook_formula():
    rng = np.random.default_rng(1)
    hues = np.concatenate([rng.uniform(0, 1, 400), [0.0, 1 / 3 - 1e-12, 1 / 3, 2 / 3, 0.999999, 0.16, 0.5, 0.631, 0.089]])
    for h in hues:
        s, i = rng.uniform(0, 1), rng.uniform(0, 1.2)
        got = po.hsi_to_rgb(float(h), float(s), float(i))
        want = hsi_to_rgb_textbook(float(h), float(s), float(i))
        np.testing.assert_allclose(got, want, rtol=0, atol=3e-15 * max(1.0, i))
        assert abs(got.mean() - i) < 1e-15 * max(1.0, i) * 4       # the I of HSI is the mean of R, G, B
    # the scene defaults: diskColor HSI(0.16, 0.1, 0.95) (src/ConfigFile.hs:73) and default.yaml's (180 deg, 0.1, 1.05)
    np.testing.assert_allclose(po.hsi_to_rgb(0.5, 0.1, 1.05), [0.945, 1.1025, 1.1025], atol=1e-12)


def test_to_word8_is_bankers_rounding_of_255x():
    L = po.lib()
    xs = np.concatenate([np.linspace(-0.2, 1.2, 2001), (np.arange(0, 256) + 0.5) / 255.0, np.arange(0, 256) / 255.0, [np.nan]])
    want = np.rint(255.0 * np.clip(np.nan_to_num(xs, nan=0.0), 0.0, 1.0)).astype(np.uint8)   # np.rint rounds half to even
    got = np.array([L.orc_to_word8(float(x)) for x in xs], dtype=np.uint8)
    np.testing.assert_array_equal(got, want)
    # exact ties in binary: 0.5/255 is not one, but 255 * (k + 0.5)/255 can be; check two by construction
    assert L.orc_to_word8(0.5 / 255 * 1.0) in (0, 1) and L.orc_to_word8(2.5 / 255.0) in (2, 3)


def look_at_rows_numpy(eye, center, up):
    """linear: lookAt eye center up = rows (xa, ya, -za) with za = normalize (center - eye),
    xa = normalize (cross za up), ya = cross xa za; normalize leaves vectors of (near) unit or zero length alone."""
    def normalize(v):
        q = float(v @ v)
        return v if (abs(q) <= 1e-12 or abs(1.0 - q) <= 1e-12) else v / np.sqrt(q)
    za = normalize(np.asarray(center, float) - np.asarray(eye, float))
    xa = normalize(np.cross(za, np.asarray(up, float)))
    return xa, np.cross(xa, za), za


def test_look_at_against_the_haddock_definition():
    import ctypes
    L = po.lib()
    rng = np.random.default_rng(2)
    for _ in range(200):
        eye, center, up = rng.normal(0, 10, 3), rng.normal(0, 3, 3), rng.normal(0, 1, 3)
        out = [np.zeros(3) for _ in range(3)]
        args = [np.ascontiguousarray(a, dtype=np.float64) for a in (eye, center, up)]
        L.orc_look_at_rows(*[a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) for a in args + out])
        for got, want in zip(out, look_at_rows_numpy(eye, center, up)):
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-14)
        xa, ya, za = out
        assert abs(xa @ ya) < 1e-14 and abs(xa @ za) < 1e-14 and abs(np.linalg.norm(za) - 1) < 1e-14


def test_in_radius_is_the_closed_ball():
    stars = starmap.synthetic_stars(30000, seed=11)
    tree = po.Tree(stars)
    rng = np.random.default_rng(3)
    pos = stars["pos"]
    for k in range(200):
        q = pos[rng.integers(len(stars))] + rng.normal(0, 0.001, 3) if k % 2 else rng.normal(0, 1, 3)
        q = q / np.linalg.norm(q)
        d = pos - q
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        want = np.sort(np.nonzero(d2 <= 0.0015 * 0.0015)[0])
        got = np.sort(tree.in_radius(0.0015, q))
        np.testing.assert_array_equal(got, want)
    # a query sitting exactly on the boundary: radius = the distance itself -> included (<=, not <)
    q = pos[0] + np.array([0.001, 0.0, 0.0])
    d = pos[0] - q
    rad = float(np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]))
    if rad * rad == (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]:
        assert 0 in tree.in_radius(rad, q)


@pytest.mark.parametrize("scene", ["default", "lensing-disk", "wideangle"])
def test_crosscheck_expectations_are_what_the_oracle_renders(scene):
    """tools/ghc_crosscheck/expected/*.png (what someone with GHC diffs the real reference against) must be
    exactly the oracle's output: regenerate three of the nine and compare the bytes."""
    from PIL import Image
    exp_dir = os.path.join(ROOT, "tools", "ghc_crosscheck", "expected")
    manifest = json.load(open(os.path.join(exp_dir, "manifest.json")))
    assert manifest["catalogue"]["sha256_ppm"] == hashlib.sha256(starmap.synthetic_catalogue()).hexdigest()
    cfg = config.prepare_scene(config.load_config(os.path.join(ROOT, "scenes", scene + ".yaml")), True)
    img, steps = po.render(cfg, po.Tree(starmap.synthetic_stars()))
    rgb8 = po.to_srgb8(img)
    meta = manifest["images"][f"prev-{scene}.png"]
    assert steps == meta["rk4_steps"] and hashlib.sha256(rgb8.tobytes()).hexdigest() == meta["sha256_rgb8"]
    np.testing.assert_array_equal(np.asarray(Image.open(os.path.join(exp_dir, f"prev-{scene}.png")).convert("RGB")), rgb8)
