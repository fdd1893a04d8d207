"""CPU tests of the KERNEL ARITHMETIC: blackstar_b200/csrc/trace_core.cuh (the code the
sm_100a kernels inline: planar RK4, MUFU-seeded |pos|^-5, disk crossing, bucketed k-d tree
lookup) is instantiated for the host by tests/hostcheck and diffed against the oracle.
This is a development aid; the parity tests proper run on the GPU (test_gpu_parity.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from blackstar_b200 import config, starmap
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-4  # north_star: per-channel max-abs error on the linear framebuffer
# The kernel arithmetic differs from the oracle's by rounding (1e-16 per operation) and by the
# force error (< 1.4e-11 relative) its |pos|^-5 primitive carries by design (trace_core.cuh:
# rinv5_seeded).  Over a few hundred steps that is ~1e-10 rad of exit direction, more for rays that skim the
# photon sphere, which the star Gaussians (sigma = 5e-4 rad) turn into ~1e-7 of colour (worst pixel seen
# on the rows through the hole of the 4096^2 headline frame: 1.1e-6).  FP64_TOL is that footprint, 50x
# inside the contract.
FP64_TOL = 2e-6


@pytest.fixture(scope="module")
def hc():
    subprocess.run(["make", "-C", os.path.join(HERE, "hostcheck")], check=True, capture_output=True)
    L = ctypes.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))
    L.hc_create.restype = ctypes.c_void_p
    L.hc_create.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    L.hc_destroy.argtypes = [ctypes.c_void_p]
    L.hc_tree_depth.argtypes = [ctypes.c_void_p]
    L.hc_render.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                            ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]
    L.hc_star_lookup.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.POINTER(ctypes.c_uint)]
    L.hc_rinv5.restype = ctypes.c_double
    L.hc_rinv5.argtypes = [ctypes.c_double]
    return L


def _hc_render(L, h, cfg, block=0):
    W, H = cfg.scene.resolution
    cam, scn = config.to_c(cfg)
    out = np.zeros((H, W, 3))
    st, hits = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    assert L.hc_render(h, ctypes.byref(cam), ctypes.byref(scn), 0, H, block, out.ctypes.data, ctypes.byref(st), ctypes.byref(hits)) == 0
    return out, st.value, hits.value


def test_rinv5_error_budget(hc):
    """0.4 q^-5/2 from a ~20-bit seed and a first-order correction: the relative error is the
    dropped 4.375 e^2 term, e = 1 - q y0^2 <= 2^-19 for the emulated seed -> <= 1.7e-11."""
    rng = np.random.default_rng(0)
    q = np.exp(rng.uniform(np.log(1e-3), np.log(1e5), 20000))
    err = max(abs(hc.hc_rinv5(float(x)) / (0.4 * x ** -2.5) - 1) for x in q)
    assert err < 4.375 * 2.0 ** -38 * 1.1


@pytest.mark.parametrize("scene", ["closeup", "default", "default-aa", "fartheraway", "lensing-disk", "lensing",
                                   "wideangle-disk", "wideangle", "wideangle1"])
def test_every_scene_matches_oracle(hc, scenes_dir, scene):
    cfg = config.load_config(f"{scenes_dir}/{scene}.yaml")
    w, h = cfg.scene.resolution
    cfg = config.with_resolution(cfg, 64, max(8, 64 * h // w))
    stars = starmap.synthetic_stars(40000, seed=5)
    tree = po.Tree(stars)
    ref, rsteps = po.render(cfg, tree)
    h_ = hc.hc_create(stars.ctypes.data, len(stars), 8)
    try:
        got, steps, _ = _hc_render(hc, h_, cfg)
        got_blocked, steps_b, _ = _hc_render(hc, h_, cfg, block=16)
    finally:
        hc.hc_destroy(h_)
    nrays = ref.shape[0] * ref.shape[1] * (4 if cfg.scene.supersampling else 1)
    assert np.abs(got - ref).max() < TOL
    # the kernel skips the reference's last (unused) RK4 evaluation of every ray
    assert steps == rsteps - nrays
    np.testing.assert_array_equal(got, got_blocked)  # step blocks do not change the arithmetic
    assert steps_b == steps


@pytest.mark.parametrize("n_stars,depth,top", [(20000, 12, 12), (100000, 14, 11), (200000, 15, 12),
                                               (468861, 16, 13), (1100000, 18, 12)])
def test_star_lookup_matches_oracle_and_brute_force(hc, n_stars, depth, top):
    # tree shapes: all levels in the shared-memory part / one and two groups of 3-level records
    hc.hc_tree_top_levels.argtypes = [ctypes.c_void_p]
    stars = starmap.synthetic_stars(n_stars, seed=17)
    tree = po.Tree(stars) if n_stars <= 200000 else None   # the restated kdt build is O(n log^2 n)
    rng = np.random.default_rng(4)
    h_ = hc.hc_create(stars.ctypes.data, len(stars), 8)
    try:
        assert hc.hc_tree_depth(h_) == depth and hc.hc_tree_top_levels(h_) == top
        nonzero = 0
        pos = stars["pos"]
        for k in range(300):
            if k % 2:
                v = pos[rng.integers(len(stars))] + rng.normal(0, 0.0007, 3)
            else:
                v = rng.normal(0, 1, 3)
            v = np.ascontiguousarray(v * rng.uniform(0.5, 2.0))  # lookup normalises (StarMap.hs:103)
            got = np.zeros(3)
            hits = ctypes.c_uint()
            hc.hc_star_lookup(h_, 0.7, 1.3, v.ctypes.data, got.ctypes.data, ctypes.byref(hits))
            n = v / np.linalg.norm(v)
            d = pos - n
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            assert hits.value == int((d2 <= 0.0015 * 0.0015).sum())
            if tree is not None:
                np.testing.assert_allclose(got, tree.lookup(0.7, 1.3, v), atol=1e-13)
            nonzero += hits.value > 0
        assert nonzero > 100
    finally:
        hc.hc_destroy(h_)


def test_tiny_and_empty_catalogues(hc, scenes_dir):
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default.yaml"), 32, 18)
    ref0, _ = po.render(cfg, None)
    for n in (0, 1, 5, 9):
        stars = starmap.synthetic_stars(n, seed=9) if n else np.zeros(0, dtype=starmap.STAR_DTYPE)
        h_ = hc.hc_create(stars.ctypes.data if n else None, n, 8)
        try:
            got, _, hits = _hc_render(hc, h_, cfg)
        finally:
            hc.hc_destroy(h_)
        ref = ref0 if n == 0 else po.render(cfg, po.Tree(stars))[0]
        assert np.abs(got - ref).max() < FP64_TOL


@pytest.mark.parametrize("case", ["cam_in_disk_plane", "cam_inside_annulus", "cam_on_y_axis", "close_small_step"])
def test_degenerate_geometry(hc, scenes_dir, case):
    import dataclasses
    base = config.load_config(f"{scenes_dir}/default.yaml")
    if case == "cam_in_disk_plane":
        cfg = config.Config(scene=dataclasses.replace(base.scene, resolution=(33, 33)),
                            camera=dataclasses.replace(base.camera, position=(0.0, 0.0, -20.0), lookAt=(0.0, 0.0, 0.0), upVec=(0.0, 1.0, 0.0)))
    elif case == "cam_inside_annulus":
        cfg = config.Config(scene=dataclasses.replace(base.scene, resolution=(32, 32), diskInner=3.0, diskOuter=30.0),
                            camera=dataclasses.replace(base.camera, position=(0.0, 0.0, -20.0), lookAt=(0.0, 0.0, 0.0), upVec=(0.0, 1.0, 0.0)))
    elif case == "cam_on_y_axis":
        cfg = config.Config(scene=dataclasses.replace(base.scene, resolution=(33, 33)),
                            camera=dataclasses.replace(base.camera, position=(0.0, 5.0, 0.0), lookAt=(0.0, 0.0, 0.0), upVec=(0.0, 0.0, 1.0)))
    else:
        cfg = config.Config(scene=dataclasses.replace(base.scene, resolution=(32, 32), stepSize=0.1),
                            camera=dataclasses.replace(base.camera, position=(3.0, 0.5, 0.0), lookAt=(0.0, 0.0, 3.0)))
    stars = starmap.synthetic_stars(20000, seed=5)
    ref, rsteps = po.render(cfg, po.Tree(stars))
    h_ = hc.hc_create(stars.ctypes.data, len(stars), 8)
    try:
        got, steps, _ = _hc_render(hc, h_, cfg)
    finally:
        hc.hc_destroy(h_)
    assert np.abs(got - ref).max() < FP64_TOL
    assert steps == rsteps - ref.shape[0] * ref.shape[1]


def _random_config(rng):
    """A random but sane scene: camera anywhere between 2.5 and 80 radii, any orientation, random
    disk, step size, field of view; sometimes exactly in the disk plane, on an axis, or inside the disk."""
    r = float(np.exp(rng.uniform(np.log(2.5), np.log(80.0))))
    d = rng.normal(0, 1, 3)
    d /= np.linalg.norm(d)
    pos = r * d
    kind = rng.integers(6)
    if kind == 0:
        pos[1] = 0.0                      # exactly in the disk plane (signum 0 at the start)
    elif kind == 1:
        pos = np.array([0.0, r, 0.0])     # on the polar axis
    look = rng.normal(0, 2.0, 3) if rng.random() < 0.7 else np.zeros(3)
    up = rng.normal(0, 1, 3)
    inner = float(rng.uniform(1.2, 4.0))
    scn = config.Scene(stepSize=float(rng.choice([0.3, 0.3, 0.2, 0.45, 0.1])), bloomStrength=0.0, bloomDivider=25,
                       starIntensity=float(rng.uniform(0.2, 1.0)), starSaturation=float(rng.uniform(0.0, 1.6)),
                       diskColor=(float(rng.uniform(0, 0.999)), float(rng.uniform(0, 0.5)), float(rng.uniform(0.5, 1.1))),
                       diskOpacity=float(rng.choice([0.0, 0.95, 0.5, 1.0])), diskInner=inner,
                       diskOuter=inner + float(rng.uniform(0.5, 20.0)), resolution=(20, 14),
                       supersampling=bool(rng.integers(2)))
    cam = config.Camera(position=tuple(float(x) for x in pos), lookAt=tuple(float(x) for x in look),
                        upVec=tuple(float(x) for x in up), fov=float(rng.uniform(0.3, 3.5)))
    return config.Config(scene=scn, camera=cam)


@pytest.mark.parametrize("seed", range(96))
def test_random_scenes_match_oracle(hc, seed):
    rng = np.random.default_rng(1000 + seed)
    cfg = _random_config(rng)
    stars = starmap.synthetic_stars(30000, seed=23)
    ref, rsteps = po.render(cfg, po.Tree(stars))
    h_ = hc.hc_create(stars.ctypes.data, len(stars), 8)
    try:
        got, steps, _ = _hc_render(hc, h_, cfg)
        got_b, steps_b, _ = _hc_render(hc, h_, cfg, block=7)      # odd block length: resumes mid-ray
    finally:
        hc.hc_destroy(h_)
    nrays = ref.shape[0] * ref.shape[1] * (4 if cfg.scene.supersampling else 1)
    assert np.isfinite(ref).all()
    assert np.abs(got - ref).max() < FP64_TOL, cfg
    assert steps == rsteps - nrays, cfg
    np.testing.assert_array_equal(got, got_b)
    assert steps_b == steps


def test_ppm_parser_equals_the_oracles_and_rejects_other_files(hc):
    """parse_ppm (what bsb_set_stars_ppm runs) against the oracle's restatement of StarMap.readMap
    (src/StarMap.hs:45-75) on the same bytes; and the refusals: a ragged length, random bytes, NaNs."""
    hc.hc_parse_ppm.restype = ctypes.c_long
    hc.hc_parse_ppm.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    data = starmap.synthetic_catalogue(50000, seed=21)
    want = po.read_ppm(data)
    out = np.zeros(50000, dtype=starmap.STAR_DTYPE)
    err = ctypes.create_string_buffer(256)
    n = hc.hc_parse_ppm(data, len(data), out.ctypes.data, len(out), err, 256)
    assert n == 50000
    for f in ("pos", "mag", "hue", "sat"):
        np.testing.assert_array_equal(out[f], want[f])
    np.testing.assert_array_equal(out["pos"], starmap.read_ppm(data)["pos"])
    # refusals (a stars.kdt or any other file used to render a garbage sky with rc = 0)
    assert hc.hc_parse_ppm(data[:-3], len(data) - 3, out.ctypes.data, len(out), err, 256) == -1 and b"PPM-format" in err.value
    rng = np.random.default_rng(5)
    junk = rng.integers(0, 256, 28 + 28 * 2000, dtype=np.uint8).tobytes()
    assert hc.hc_parse_ppm(junk, len(junk), out.ctypes.data, len(out), err, 256) == -1 and b"not a star" in err.value
    assert hc.hc_parse_ppm(b"tiny", 4, out.ctypes.data, len(out), err, 256) == -1
    # the flat-list entry point validates too
    hc.hc_validate_stars.restype = ctypes.c_char_p
    hc.hc_validate_stars.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    assert hc.hc_validate_stars(out.ctypes.data, len(out)) == b""
    bad = out.copy()
    bad["pos"][7] = [np.nan, 0, 0]
    assert b"star 7" in hc.hc_validate_stars(bad.ctypes.data, len(bad))
    bad = out.copy()
    bad["hue"][9] = 1.25
    assert b"star 9" in hc.hc_validate_stars(bad.ctypes.data, len(bad))


def test_star_tree_build_time(hc):
    """N3: the host-side tree build that replaces `generate-tree` (src/StarMap.hs:90-91) on the full catalogue."""
    hc.hc_build_tree_ms.restype = ctypes.c_double
    hc.hc_build_tree_ms.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    stars = starmap.synthetic_stars()
    ms = hc.hc_build_tree_ms(stars.ctypes.data, len(stars), 3)
    print(f"build_star_tree({len(stars)} stars): {ms:.1f} ms on this host")
    assert ms < 1500


def test_stars_kdt_tree_file(hc):
    """The reference's own star map file (stars.kdt, what --starmap defaults to, app/Main.hs:36): a tree file
    written in the (recalled) cereal/kdt layout decodes to the same star set as the PPM catalogue it was made
    from -- and anything that violates one of the layout's invariants is refused, not rendered."""
    hc.hc_parse_star_file.restype = ctypes.c_long
    hc.hc_parse_star_file.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    ppm = starmap.synthetic_catalogue(20000, seed=33)
    want = starmap.read_ppm(ppm)
    kdt = starmap.catalogue_to_kdt(ppm)
    assert kdt[:3] == b"\x00\x00\x00" and len(kdt) == 2 + 20000 * (1 + 24 + 8 + 1 + 8) + 20001 + 8
    out = np.zeros(20000, dtype=starmap.STAR_DTYPE)
    err = ctypes.create_string_buffer(512)
    assert hc.hc_parse_star_file(kdt, len(kdt), out.ctypes.data, len(out), err, 512) == 20000, err.value
    key = lambda a: np.lexsort((a["pos"][:, 2], a["pos"][:, 1], a["pos"][:, 0]))
    got, ref = out[key(out)], want[key(want)]
    for f in ("pos", "mag", "hue", "sat"):
        np.testing.assert_array_equal(got[f], ref[f])
    # the same entry point still takes the catalogue itself
    assert hc.hc_parse_star_file(ppm, len(ppm), out.ctypes.data, len(out), err, 512) == 20000
    # refusals: truncated, a flipped tag, a perturbed coordinate (breaks axisValue == coordinate or the unit norm),
    # a wrong size field, trailing bytes
    bad = bytearray(kdt)
    assert hc.hc_parse_star_file(bytes(bad[:-9]), len(bad) - 9, out.ctypes.data, len(out), err, 512) == -1
    bad = bytearray(kdt); bad[2] = 7
    assert hc.hc_parse_star_file(bytes(bad), len(bad), out.ctypes.data, len(out), err, 512) == -1
    i = kdt.index(b"\x01\x3f") if b"\x01\x3f" in kdt else 40
    bad = bytearray(kdt); bad[len(bad) // 2] ^= 0x10
    assert hc.hc_parse_star_file(bytes(bad), len(bad), out.ctypes.data, len(out), err, 512) == -1
    bad = bytearray(kdt); bad[-1] ^= 1
    assert hc.hc_parse_star_file(bytes(bad), len(bad), out.ctypes.data, len(out), err, 512) == -1 and b"size" in err.value
    bad = bytearray(kdt) + b"\x00"
    assert hc.hc_parse_star_file(bytes(bad), len(bad), out.ctypes.data, len(out), err, 512) == -1
    junk = np.random.default_rng(1).integers(0, 256, 5000, dtype=np.uint8).tobytes()
    assert hc.hc_parse_star_file(junk, len(junk), out.ctypes.data, len(out), err, 512) == -1 and b"neither" in err.value
