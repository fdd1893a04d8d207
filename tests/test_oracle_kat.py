"""Known-answer tests that PIN THE ORACLE.

The reference ships no tests and cannot be built here, so the oracle is pinned against
closed-form facts about the algorithm of src/Raytracer.hs / src/ImageFilters.hs that do
not depend on any un-shipped source.
"""
import dataclasses
import math

import numpy as np
import pytest

from blackstar_b200 import config
from oracle import pyoracle as po


@pytest.fixture(scope="module")
def default_cfg(scenes_dir):
    return config.load_config(f"{scenes_dir}/default.yaml")


def test_hsi_mean_is_intensity_and_disk_colour():
    # HSI -> RGB keeps mean(R,G,B) = I by construction; default disk colour KAT
    rgb = po.hsi_to_rgb(0.5, 0.1, 1.05)
    np.testing.assert_allclose(rgb, [0.945, 1.1025, 1.1025], rtol=0, atol=1e-15)
    rng = np.random.default_rng(0)
    for _ in range(200):
        h, s, i = rng.uniform(0, 0.999999), rng.uniform(0, 1.5), rng.uniform(0, 1.2)
        assert abs(po.hsi_to_rgb(h, s, i).mean() - i) < 1e-14
    assert np.isnan(po.hsi_to_rgb(1.0, 0.5, 0.5)).all()   # h >= 2 pi is an `error` in massiv-io
    assert np.isnan(po.hsi_to_rgb(-0.1, 0.5, 0.5)).all()


def test_rk4_conserves_energy_and_angular_momentum():
    # E = |v|^2/2 - h2/(2 r^3) and |pos x vel|^2 are invariants of f (src/Raytracer.hs:126-127)
    pos = np.array([0.0, 1.0, -20.0])
    for b, tol in ((5.0, 1e-7), (8.0, 1e-7), (2.7, 4e-5)):
        # inward ray with impact parameter b
        r0 = np.linalg.norm(pos)
        e1 = pos / r0
        e2 = np.array([1.0, 0.0, 0.0])
        s = b / r0
        vel = -math.sqrt(1 - s * s) * e1 + s * e2
        h2 = float(np.dot(np.cross(pos, vel), np.cross(pos, vel)))
        E0 = 0.5 * vel @ vel - h2 / (2 * r0 ** 3)
        p, v = pos.copy(), vel.copy()
        for _ in range(400):
            v, p = po.rk4(0.3, h2, v, p)
            r = np.linalg.norm(p)
            if r < 1 or r > 60:
                break
            L2 = np.cross(p, v) @ np.cross(p, v)
            E = 0.5 * v @ v - h2 / (2 * r ** 3)
            assert abs(L2 - h2) <= tol * h2
            assert abs(E - E0) <= tol


def test_capture_threshold(default_cfg):
    # from r0 = sqrt(401) an inward ray is captured iff h^2 < 1/(4/27 + 1/r0^3): h_c = 2.59698;
    # RK4 at step 0.3 agrees to ~2e-5
    r0 = math.sqrt(401.0)
    hc = 1 / math.sqrt(4 / 27 + 1 / r0 ** 3)
    assert abs(hc - 2.59698) < 1e-5
    pos = np.array([0.0, 1.0, -20.0])
    e1 = pos / r0
    e2 = np.array([1.0, 0.0, 0.0])

    def captured(b):
        s = b / r0
        vel = -math.sqrt(1 - s * s) * e1 + s * e2
        h2 = float(np.dot(np.cross(pos, vel), np.cross(pos, vel)))
        p, v = pos.copy(), vel.copy()
        for _ in range(5000):
            v, p = po.rk4(0.3, h2, v, p)
            r2 = p @ p
            if r2 < 1:
                return True
            if r2 > 2500:
                return False
        raise AssertionError("orbiting")

    lo, hi = 2.0, 3.2
    assert captured(lo) and not captured(hi)
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        if captured(mid):
            lo = mid
        else:
            hi = mid
    assert abs(lo - hc) < 1e-4


def test_radial_ray_is_straight_and_captured(default_cfg):
    cam = dataclasses.replace(default_cfg.camera, lookAt=(0.0, 0.0, 0.0))
    cfg = config.Config(scene=dataclasses.replace(default_cfg.scene, resolution=(2, 2)), camera=cam)
    vel, pos = po.generate_ray(cfg, 2, 2, 1, 1)  # pixel (w/2, h/2) looks straight down the axis
    r0 = np.linalg.norm(pos)
    np.testing.assert_allclose(vel, -pos / r0, atol=1e-15)
    rgb, steps = po.trace_ray(cfg, None, 2, 2, 1, 1)
    assert (rgb == 0).all()
    # straight line at unit speed: r < 1 after ceil((r0 - 1)/0.3) = 64 steps, +1 rk4 evaluated before the test
    assert steps == math.ceil((r0 - 1) / 0.3) + 1


def test_step_count_default_scene(default_cfg):
    # SURVEY.md 4.2 / BASELINE.md 3: 28 968 716 steps over 129 600 rays at 480x270
    cfg = config.with_resolution(default_cfg, 480, 270)
    img, steps = po.render(cfg)
    assert steps == 28968716
    assert img.shape == (270, 480, 3)
    assert img.min() >= 0 and img.max() < 1.2


def test_generate_ray_pixel_corner_convention(default_cfg):
    # vx = fov*(x/w - .5), vy = fov*(.5 - y/h)*h/w, dir = normalize(xa vx + ya vy + za) (Raytracer.hs:47-51)
    w, h = 64, 36
    cam = default_cfg.camera
    pos, look, up = map(np.array, (cam.position, cam.lookAt, cam.upVec))
    za = (look - pos) / np.linalg.norm(look - pos)
    xa = np.cross(za, up); xa /= np.linalg.norm(xa)
    ya = np.cross(xa, za)
    for (x, y) in ((0, 0), (63, 35), (32, 18), (5, 30)):
        v, p = po.generate_ray(default_cfg, w, h, x, y)
        d = xa * (cam.fov * (x / w - 0.5)) + ya * (cam.fov * (0.5 - y / h) * h / w) + za
        np.testing.assert_allclose(v, d / np.linalg.norm(d), atol=2e-16)
        np.testing.assert_array_equal(p, pos)


def test_supersample_is_exact_block_mean():
    rng = np.random.default_rng(1)
    img = rng.uniform(0, 1, (6, 10, 3))
    out = po.supersample(img)
    ref = 0.25 * (((img[0::2, 0::2] + img[1::2, 0::2]) + img[0::2, 1::2]) + img[1::2, 1::2])
    np.testing.assert_array_equal(out, ref)


def test_box_blur_impulse_and_dc_gain():
    # one 1-D pass: impulse at x0 -> 1/(2r+1) on [x0-r, x0+r-1] (window [x-r+1, x+r], divisor 2r+1)
    r, n = 3, 32
    img = np.zeros((1, n, 3))
    img[0, 16] = 1.0
    out = po.box_blur(r, 1, img)  # H then V; V on a 1-row image: window covers only row 0 -> 1/(2r+1)
    expect = np.zeros(n)
    expect[16 - r:16 + r] = 1.0 / (2 * r + 1)
    np.testing.assert_allclose(out[0, :, 0], expect / (2 * r + 1), atol=1e-17)
    # DC gain (2r/(2r+1)) per 1-D pass in the interior
    flat = np.ones((64, 64, 3))
    b = po.box_blur(2, 1, flat)
    np.testing.assert_allclose(b[32, 32], (4 / 5) ** 2, atol=1e-15)
    b3 = po.box_blur(2, 3, flat)
    np.testing.assert_allclose(b3[32, 32], (4 / 5) ** 6, atol=1e-14)


def test_box_blur_matches_literal_numpy_restatement():
    rng = np.random.default_rng(2)
    img = rng.uniform(0, 1, (17, 23, 3))
    r = 4

    def pass1d(a, axis):
        a = np.moveaxis(a, axis, 0)
        n = a.shape[0]
        pad = np.zeros((r + 1,) + a.shape[1:])
        ap = np.concatenate([pad, a, pad], axis=0)
        out = np.zeros_like(a)
        for x in range(n):
            out[x] = ap[x + r + 1 - r + 1: x + r + 1 + r + 1].sum(axis=0) / (2 * r + 1)
        return np.moveaxis(out, 0, axis)

    ref = img.copy()
    for _ in range(3):
        ref = pass1d(ref, 1)
        ref = pass1d(ref, 0)
    np.testing.assert_allclose(po.box_blur(r, 3, img), ref, atol=1e-14)
    np.testing.assert_allclose(po.bloom(0.15, 5, img), img + 0.15 * ref, atol=1e-14)  # r = 23 // 5 = 4


def test_srgb_and_word8():
    L = po.lib()
    assert L.orc_srgb(0.0) == 0.0
    assert abs(L.orc_srgb(0.0031308) - 12.92 * 0.0031308) < 1e-6   # continuity at the knee
    assert abs(L.orc_srgb(1.0) - 1.0) < 1e-15
    assert L.orc_to_word8(0.5 / 255) == 0      # half-to-even: 0.5 -> 0
    assert L.orc_to_word8(1.5 / 255) == 2      # 1.5 -> 2
    assert L.orc_to_word8(2.5 / 255) == 2      # 2.5 -> 2
    assert L.orc_to_word8(-3.0) == 0 and L.orc_to_word8(7.0) == 255


def test_in_radius_matches_brute_force(small_stars):
    tree = po.Tree(small_stars)
    rng = np.random.default_rng(3)
    pos = small_stars["pos"]
    for k in range(300):
        if k % 2:
            q = pos[rng.integers(len(pos))] + rng.normal(0, 0.004, 3)
        else:
            q = rng.normal(0, 1, 3)
        q /= np.linalg.norm(q)
        for rad in (0.0015, 0.02):
            got = np.sort(tree.in_radius(rad, q))
            d = pos - q
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            want = np.flatnonzero(d2 <= rad * rad)
            np.testing.assert_array_equal(got, want)


def test_ppm_reader_matches_host_mirror():
    from blackstar_b200 import starmap
    data = starmap.synthetic_catalogue(5000, seed=11)
    a = po.read_ppm(data)
    b = starmap.read_ppm(data)
    assert len(a) == len(b) == 5000
    np.testing.assert_array_equal(a["mag"], b["mag"])
    np.testing.assert_array_equal(a["hue"], b["hue"])
    np.testing.assert_array_equal(a["sat"], b["sat"])
    np.testing.assert_allclose(a["pos"], b["pos"], atol=3e-16)
    np.testing.assert_allclose(np.linalg.norm(a["pos"], axis=1), 1.0, atol=1e-15)
