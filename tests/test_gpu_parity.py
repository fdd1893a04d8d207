"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle and the
committed golden fixtures.  Tolerance (north_star): per-channel max-abs error < 1e-4 on the
linear float framebuffer; integer/byte outputs (sRGB8) bit-exact."""
import ctypes
import dataclasses
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

from blackstar_b200 import _lib, config, starmap  # noqa: E402
from blackstar_b200.render import Renderer  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

TOL = 1e-4
SCENES = ["closeup", "default", "default-aa", "fartheraway", "lensing-disk", "lensing", "wideangle-disk",
          "wideangle", "wideangle1"]


@pytest.fixture(scope="module")
def rnd():
    r = Renderer(devices=[0])   # raises (no fallback) if the .so or the GPU is missing
    yield r
    r.close()


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "scenes_48.npz"))


@pytest.fixture(scope="module")
def stars40k():
    return starmap.synthetic_stars(40000, seed=5)


def rgb(img):
    return img[..., :3].astype(np.float64)


def test_rinv5_primitive_on_device(rnd):
    rel, seed = rnd.selftest_rinv5(0.25, 1e5, 1 << 20)
    print(f"rinv5k max rel err {rel:.3e}; MUFU.RSQ64H seed residual {seed:.3e} (2^{np.log2(seed):.1f})")
    assert seed < 2.0 ** -18.5                  # MUFU.RSQ64H: |e| = |1 - q y0^2| <~ 2^-19.1
    assert rel < 4.375 * seed * seed * 1.05 + 1e-15   # the dropped second-order term, nothing else
    assert rel < 3e-11


@pytest.mark.parametrize("scene", SCENES)
def test_golden_scenes(rnd, golden, scenes_dir, scene):
    cfg = make_golden.golden_config(f"{scenes_dir}/{scene}.yaml")
    rnd.set_stars(starmap.synthetic_stars(**make_golden.GOLDEN_STARS))
    rnd.set_option("trace_variant", 0)
    img = rnd.render(cfg)
    st = rnd.last_stats
    nrays = img.shape[0] * img.shape[1] * (4 if cfg.scene.supersampling else 1)
    assert np.abs(rgb(img) - golden[scene + "/render"]).max() < TOL
    assert (img[..., 3] == 1).all()
    assert st["rays"] == nrays and st["capped"] == 0
    assert st["steps"] == int(golden[scene + "/steps"][0]) - nrays  # the unused last RK4 of each ray is skipped
    full = rnd.do_render(cfg)
    assert np.abs(rgb(full) - golden[scene + "/bloomed"]).max() < TOL
    u8 = rnd.do_render_srgb8(cfg)
    diff = np.abs(u8.astype(int) - golden[scene + "/srgb8"].astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 2e-3  # float32 framebuffer vs f64: rare 1-LSB flips


@pytest.mark.parametrize("scene", SCENES)
def test_scene_vs_oracle_all_schedules(rnd, scenes_dir, stars40k, scene):
    cfg = config.load_config(f"{scenes_dir}/{scene}.yaml")
    w, h = cfg.scene.resolution
    cfg = config.with_resolution(cfg, 136, 136 * h // w + 1)   # odd sizes exercise tile padding
    rnd.set_stars(stars40k)
    ref, rsteps = po.render(cfg, po.Tree(stars40k))
    imgs = []
    for variant in (0, 1, 2, 3, 4, 6):
        rnd.set_option("trace_variant", variant)
        img = rnd.render(cfg)
        assert np.abs(rgb(img) - ref).max() < TOL, f"variant {variant}"
        nrays = img.shape[0] * img.shape[1] * (4 if cfg.scene.supersampling else 1)
        assert rnd.last_stats["steps"] == rsteps - nrays, f"variant {variant}"
        imgs.append(img)
    rnd.set_option("trace_variant", 0)
    for im in imgs[1:]:
        np.testing.assert_array_equal(im, imgs[0])  # schedules differ, arithmetic does not


def test_row_tiles_concatenate_bit_exactly(rnd, scenes_dir, stars40k):
    # multi-GPU correctness without a cluster: N row-tile renders == 1 render
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default-aa.yaml"), 200, 113)
    rnd.set_stars(stars40k)
    whole = rnd.render(cfg)
    for n in (2, 3, 8):
        H = 113
        parts = [rnd.render(cfg, H * k // n, H * (k + 1) // n) for k in range(n)]
        np.testing.assert_array_equal(np.concatenate(parts, axis=0), whole)
    assert rnd.render(cfg, 5, 5).shape == (0, 200, 4)  # empty tile is legal


def test_empty_star_map_is_black_sky(rnd, scenes_dir):
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default.yaml"), 160, 90)
    rnd.set_stars(None)
    assert rnd.star_count == 0
    img = rnd.render(cfg)
    ref, _ = po.render(cfg, None)
    assert np.abs(rgb(img) - ref).max() < TOL
    assert rnd.last_stats["star_hits"] == 0


def test_ppm_ingestion_equals_flat_list(rnd, scenes_dir):
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/wideangle1.yaml"), 128, 72)
    data = starmap.synthetic_catalogue(50000, seed=21)
    rnd.set_stars_ppm(data)
    assert rnd.star_count == 50000
    a = rnd.render(cfg)
    rnd.set_stars(starmap.read_ppm(data))
    b = rnd.render(cfg)
    assert np.abs(a - b).max() < 1e-6   # libm vs numpy cos/sin may differ by an ulp in star positions
    with pytest.raises(_lib.BlackstarError):
        rnd.set_stars_ppm(b"tiny")


@pytest.mark.parametrize("shape,divider", [((40, 64), 25), ((64, 40), 7), ((300, 500), 25), ((270, 480), 3),
                                           ((90, 1300), 25), ((1100, 70), 2), ((33, 2500), 25), ((17, 4100), 25),
                                           ((4100, 26), 25), ((37, 53), 5), ((1, 30), 3), ((30, 1), 1),
                                           ((255, 1023), 9), ((512, 2048), 25)])
def test_bloom_vs_oracle(rnd, shape, divider):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    H, W = shape
    img = np.zeros((H, W, 4), dtype=np.float32)
    img[..., :3] = rng.uniform(0, 1.2, (H, W, 3)).astype(np.float32)
    img[..., 3] = 1
    img[rng.integers(H), rng.integers(W), :3] = 50.0  # a hot pixel (stars/disk produce these)
    got = rnd.bloom(0.4, divider, img)
    ref = po.bloom(0.4, divider, rgb(img))
    assert np.abs(rgb(got) - ref).max() < 1e-5
    assert (got[..., 3] == 1).all()


def test_stars_kdt_ingestion_equals_flat_list(rnd, scenes_dir):
    # --starmap's own format: a tree file in the (recalled) stars.kdt layout renders the same frame as the flat list
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default.yaml"), 160, 90)
    ppm = starmap.synthetic_catalogue(30000, seed=9)
    rnd.set_stars_file(starmap.catalogue_to_kdt(ppm))
    assert rnd.star_count == 30000
    a = rnd.render(cfg)
    rnd.set_stars_file(ppm)
    b = rnd.render(cfg)
    np.testing.assert_array_equal(a, b)     # same star set -> same tree -> same bits
    with pytest.raises(_lib.BlackstarError) as e:
        rnd.set_stars_file(b"\x00\x00\x05 not a tree")
    assert "neither" in e.value.message


def test_bloom_full_frame_4096_vs_oracle(rnd, scenes_dir):
    """The WHOLE bloomed headline frame against the oracle (ImageFilters.hs:28-86): the rendered
    4096x4096 default-aa frame (stars + disk: point-like maxima next to black) goes through the
    GPU bloom and through the C restatement of boxBlur's sequential running sums."""
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default-aa.yaml"), 4096, 4096)
    rnd.set_stars(starmap.synthetic_stars())
    pre = rnd.render(cfg)
    got = rnd.bloom(cfg.scene.bloomStrength, cfg.scene.bloomDivider, pre)
    ref = po.bloom(cfg.scene.bloomStrength, cfg.scene.bloomDivider, rgb(pre))
    err = np.abs(rgb(got) - ref)
    print(f"4096^2 bloom vs oracle: max {err.max():.3e}, mean {err.mean():.3e}; max(ref) = {ref.max():.3f}")
    assert err.max() < 1e-5
    assert (got[..., 3] == 1).all()
    # doRender = render + bloom in one call lands on the same pixels, bit for bit
    np.testing.assert_array_equal(rnd.do_render(cfg), got)
    # ... and so does the fused sRGB8 epilogue: it is the 8-bit map of exactly those floats
    np.testing.assert_array_equal(rnd.do_render_srgb8(cfg), po.to_srgb8(rgb(got)))


@pytest.mark.parametrize("shape,divider", [((24, 9000), 25), ((8300, 20), 2), ((3, 16500), 40)])
def test_bloom_lines_longer_than_the_shared_memory_kernel(rnd, shape, divider):
    # the reference's boxBlur has no size limit: sides above 8192 take the sequential running-sum path
    rng = np.random.default_rng(shape[0])
    H, W = shape
    img = np.ones((H, W, 4), dtype=np.float32)
    img[..., :3] = rng.uniform(0, 1.2, (H, W, 3)).astype(np.float32)
    got = rnd.bloom(0.4, divider, img)
    ref = po.bloom(0.4, divider, rgb(img))
    assert np.abs(rgb(got) - ref).max() < 1e-5
    assert (got[..., 3] == 1).all()


@pytest.mark.parametrize("res", [(333, 187), (640, 360), (2048, 30)])
def test_fused_srgb8_epilogue_equals_separate_map(rnd, scenes_dir, stars40k, res):
    # bsb_render_full_srgb8 (sRGB + toWord8 in the epilogue of the second bloom launch) against
    # bsb_render_full followed by the oracle's writeImg map; odd widths take the byte-store path
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default-aa.yaml"), *res)
    rnd.set_stars(stars40k)
    full = rnd.do_render(cfg)
    np.testing.assert_array_equal(rnd.do_render_srgb8(cfg), po.to_srgb8(rgb(full)))
    nobloom = config.Config(scene=dataclasses.replace(cfg.scene, bloomStrength=0.0), camera=cfg.camera)
    np.testing.assert_array_equal(rnd.do_render_srgb8(nobloom), po.to_srgb8(rgb(rnd.do_render(nobloom))))


@pytest.mark.parametrize("shape,divider,world", [((96, 200), 10, 3), ((1080, 1920), 25, 8), ((130, 70), 4, 2), ((64, 2100), 25, 4)])
def test_distributed_bloom_building_blocks_on_one_gpu(rnd, shape, divider, world):
    """bsb_bloom_h_device / bsb_bloom_v_device as the N-rank pipeline uses them, with the all-to-all
    done by hand on one GPU: row tiles -> H^3 (transposed) -> column bands from row-tile pieces -> V^3 +
    combine + sRGB8.  Against the oracle, the one-piece bloom of the library and the oracle's 8-bit map."""
    import torch
    from blackstar_b200.dist import col_bands, even_row_tiles
    H, W = shape
    rng = np.random.default_rng(H + W)
    img = np.ones((H, W, 4), dtype=np.float32)
    img[..., :3] = rng.uniform(0, 1.1, (H, W, 3)).astype(np.float32)
    ref = po.bloom(0.35, divider, rgb(img))
    one = rnd.bloom(0.35, divider, img)
    r = W // divider
    tiles, bands = even_row_tiles(H, world), col_bands(W, world)
    rnd.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        dev = torch.from_numpy(img).cuda()
        midT, imgT = [], []
        for r0, r1 in tiles:
            m = torch.empty((W, r1 - r0, 4), device="cuda")
            t = torch.empty((W, r1 - r0, 4), device="cuda")
            rnd.bloom_h_device(r, W, r1 - r0, dev[r0:r1].data_ptr(), m.data_ptr(), t.data_ptr())
            midT.append(m)
            imgT.append(t)
        torch.cuda.synchronize()
        for (r0, r1), t in zip(tiles, imgT):       # the transposed copy of the tile is exact
            assert torch.equal(t, dev[r0:r1].permute(1, 0, 2))
        out = np.zeros((H, W, 4), dtype=np.float32)
        out8 = np.zeros((H, W, 3), dtype=np.uint8)
        for c0, c1 in bands:
            if c1 == c0:
                continue
            pm = [m[c0:c1].contiguous() for m in midT]     # what the all-to-all delivers: one piece per source rank
            pi = [t[c0:c1].contiguous() for t in imgT]
            band = torch.empty((H, c1 - c0, 4), device="cuda")
            band8 = torch.empty((H, c1 - c0, 3), device="cuda", dtype=torch.uint8)
            rnd.bloom_v_device(0.35, r, H, c1 - c0, [x.data_ptr() for x in pm], [x.data_ptr() for x in pi],
                               [b - a for a, b in tiles], band.data_ptr(), band8.data_ptr())
            torch.cuda.synchronize()
            out[:, c0:c1] = band.cpu().numpy()
            out8[:, c0:c1] = band8.cpu().numpy()
    finally:
        rnd.set_stream(None)
    assert np.abs(rgb(out) - ref).max() < 1e-5
    assert np.abs(out - one).max() < 2e-6          # same filter, different chunking of the prefix sums
    assert (out[..., 3] == 1).all()
    np.testing.assert_array_equal(out8, po.to_srgb8(rgb(out)))


def test_bloom_device_in_place_and_unaligned_views(rnd):
    # the device entry point on torch tensors: in place, out of place, and on a view whose base
    # address is only 16-byte aligned (the 256-bit path must not be taken there)
    import torch
    rng = np.random.default_rng(9)
    H, W = 96, 200
    img = np.ones((H, W, 4), dtype=np.float32)
    img[..., :3] = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    ref = po.bloom(0.3, 10, rgb(img))
    rnd.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        a = torch.from_numpy(img).cuda()
        b = torch.empty_like(a)
        rnd.bloom_device(0.3, 10, W, H, a.data_ptr(), b.data_ptr())
        rnd.bloom_device(0.3, 10, W, H, a.data_ptr(), a.data_ptr())
        big = torch.zeros(H * W * 4 + 4, dtype=torch.float32, device="cuda")
        view = big[4:]                                 # +16 bytes
        view.copy_(torch.from_numpy(img).cuda().reshape(-1))
        rnd.bloom_device(0.3, 10, W, H, view.data_ptr(), view.data_ptr())
        torch.cuda.synchronize()
        for got in (b.cpu().numpy(), a.cpu().numpy(), view.cpu().numpy().reshape(H, W, 4)):
            assert np.abs(got[..., :3].astype(np.float64) - ref).max() < 1e-5
    finally:
        rnd.set_stream(None)


def test_bloom_error_behaviour(rnd):
    img = np.ones((8, 10, 4), dtype=np.float32)
    with pytest.raises(_lib.BlackstarError) as e:
        rnd.bloom(0.4, 25, img)     # r = 10 div 25 = 0: the reference's boxBlur crashes (foldl1' [])
    assert e.value.code == 1
    with pytest.raises(_lib.BlackstarError):
        rnd.bloom(0.4, 0, img)


def test_srgb8_bit_exact(rnd):
    rng = np.random.default_rng(5)
    img = np.ones((37, 53, 4), dtype=np.float32)
    img[..., :3] = rng.uniform(-0.1, 1.3, (37, 53, 3)).astype(np.float32)
    img[0, 0, :3] = [0.0, 0.0031308, 1.0]
    got = rnd.to_srgb8(img)
    ref = po.to_srgb8(rgb(img))
    np.testing.assert_array_equal(got, ref)


def test_srgb8_bit_exact_around_every_threshold(rnd):
    """The device map is a threshold table + a MUFU guess; check it where it could go wrong: the two
    floats on either side of every one of the 255 level changes, the sRGB knee, and a dense sweep."""
    xs = [np.float32(0.0), np.float32(0.0031308), np.float32(1.0), np.float32(-1.0), np.float32(7.5), np.float32(np.nan)]
    lo, hi = np.float32(0.0), np.float32(1.0)
    grid = np.linspace(0, 1, 1 << 16, dtype=np.float32)
    lev = po.to_srgb8(np.repeat(grid[:, None], 3, 1).astype(np.float64)[None])[0, :, 0]
    for k in np.nonzero(np.diff(lev.astype(int)))[0]:
        a, b = grid[k], grid[k + 1]                 # level changes somewhere in (a, b]
        while np.nextafter(a, b) < b:
            m = np.float32((np.float64(a) + np.float64(b)) / 2)
            if m == a or m == b:
                break
            lm = po.to_srgb8(np.full((1, 1, 3), np.float64(m)))[0, 0, 0]
            if lm == lev[k]:
                a = m
            else:
                b = m
        xs += [np.nextafter(a, lo), a, b, np.nextafter(b, hi)]
    rng = np.random.default_rng(3)
    xs = np.concatenate([np.array(xs, dtype=np.float32), rng.uniform(0, 1.02, 300000).astype(np.float32),
                         np.exp(rng.uniform(np.log(1e-8), 0, 100000)).astype(np.float32)])
    n = (len(xs) + 2) // 3 * 3
    xs = np.resize(xs, n)
    img = np.ones((1, n // 3, 4), dtype=np.float32)
    img[0, :, :3] = xs.reshape(-1, 3)
    got = rnd.to_srgb8(img)
    ref = po.to_srgb8(rgb(img))
    np.testing.assert_array_equal(got, ref)
    assert len(set(ref.reshape(-1).tolist())) == 256


def test_pageable_and_pinned_host_buffers_give_the_same_frame(rnd, scenes_dir, stars40k):
    # bsb_render_full into a page-locked buffer (one DMA) and into plain malloc memory (the staged,
    # multi-threaded copy): 1920x1080x16 B = 33 MB = several staging chunks with a ragged last one
    import torch
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default.yaml"), 1920, 1081)
    rnd.set_stars(stars40k)
    pinned = torch.empty((1081, 1920, 4), dtype=torch.float32, pin_memory=True)
    a = rnd.do_render(cfg, out=pinned.numpy())
    b = rnd.do_render(cfg, out=np.full((1081, 1920, 4), np.nan, dtype=np.float32))
    np.testing.assert_array_equal(a, b)
    p8 = torch.empty((1081, 1920, 3), dtype=torch.uint8, pin_memory=True)
    np.testing.assert_array_equal(rnd.do_render_srgb8(cfg, out=p8.numpy()), rnd.do_render_srgb8(cfg))


@pytest.mark.parametrize("scene", SCENES)
def test_ghc_crosscheck_expectations(rnd, scenes_dir, scene):
    """The CUDA path renders the committed cross-check images (tools/ghc_crosscheck/expected: what the real
    reference is to be diffed against by someone with GHC): `--preview` of every scene, RGB8."""
    from PIL import Image
    exp = np.asarray(Image.open(os.path.join(HERE, "..", "tools", "ghc_crosscheck", "expected", f"prev-{scene}.png")).convert("RGB"))
    cfg = config.prepare_scene(config.load_config(f"{scenes_dir}/{scene}.yaml"), True)
    rnd.set_stars(starmap.synthetic_stars())
    got = rnd.do_render_srgb8(cfg)
    assert got.shape == exp.shape
    d = np.abs(got.astype(int) - exp.astype(int))
    assert d.max() <= 1 and (d.max(axis=2) > 0).mean() < 2e-3   # float32 framebuffer vs the oracle's f64: rare 1-LSB flips


def test_step_cap_is_an_option_and_is_reported_on_every_path(rnd, scenes_dir):
    """The reference has no iteration cap (src/Raytracer.hs:80-85); ours is a safety net the caller can move
    ("step_cap").  A capped ray must surface as BSB_ERR_STEPCAP on the synchronous calls AND, through
    bsb_synchronize, on the asynchronous ones (the async path used to drop it)."""
    import torch
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default.yaml"), 64, 36)
    rnd.set_stars(None)
    ok = rnd.render(cfg)
    rnd.set_option("step_cap", 100)            # every ray of this scene needs > 200 steps
    try:
        with pytest.raises(_lib.BlackstarError) as e:
            rnd.render(cfg)
        assert e.value.code == 5 and rnd.last_stats["capped"] > 1000   # the rays that fall in within 100 steps are not capped
        with pytest.raises(_lib.BlackstarError) as e:
            rnd.do_render(cfg)
        assert e.value.code == 5
        buf = torch.empty((36, 64, 4), device="cuda")
        rnd.render_device(cfg, buf.data_ptr())               # asynchronous: returns before the kernel has run
        with pytest.raises(_lib.BlackstarError) as e:
            rnd.synchronize()
        assert e.value.code == 5
    finally:
        rnd.set_option("step_cap", 1000000)
    np.testing.assert_array_equal(rnd.render(cfg), ok)
    rnd.synchronize()
    with pytest.raises(_lib.BlackstarError):
        rnd.set_option("step_cap", 0)


def test_invalid_arguments_return_status_not_crash(rnd, scenes_dir):
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default.yaml"), 32, 18)
    with pytest.raises(_lib.BlackstarError):
        rnd.render(cfg, 3, 99)
    bad = config.Config(scene=dataclasses.replace(cfg.scene, stepSize=0.0), camera=cfg.camera)
    with pytest.raises(_lib.BlackstarError):
        rnd.render(bad)
    bad = config.Config(scene=cfg.scene, camera=dataclasses.replace(cfg.camera, fov=float("nan")))
    with pytest.raises(_lib.BlackstarError):
        rnd.render(bad)
    with pytest.raises(_lib.BlackstarError):
        rnd.set_option("no_such_option", 1)
    # the ctx is still usable afterwards
    assert rnd.render(cfg).shape == (18, 32, 4)


def test_camera_inside_horizon_and_radial_ray(rnd, scenes_dir):
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default.yaml"), 16, 16)
    rnd.set_stars(None)
    inside = config.Config(scene=cfg.scene, camera=dataclasses.replace(cfg.camera, position=(0.3, 0.2, 0.1)))
    assert (rnd.render(inside)[..., :3] == 0).all()
    radial = config.Config(scene=dataclasses.replace(cfg.scene, resolution=(2, 2)),
                           camera=dataclasses.replace(cfg.camera, lookAt=(0.0, 0.0, 0.0)))
    img = rnd.render(radial)
    ref, _ = po.render(radial, None)
    assert np.abs(rgb(img) - ref).max() < TOL


@pytest.mark.parametrize("case", ["cam_in_disk_plane", "cam_inside_annulus", "cam_on_y_axis", "close_small_step"])
def test_degenerate_geometry(rnd, scenes_dir, stars40k, case):
    # orbital planes through the y axis / in the disk plane, exact-zero signum at the start
    base = config.load_config(f"{scenes_dir}/default.yaml")
    if case == "cam_in_disk_plane":
        cfg = config.Config(scene=dataclasses.replace(base.scene, resolution=(65, 65)),
                            camera=dataclasses.replace(base.camera, position=(0.0, 0.0, -20.0), lookAt=(0.0, 0.0, 0.0), upVec=(0.0, 1.0, 0.0)))
    elif case == "cam_inside_annulus":
        cfg = config.Config(scene=dataclasses.replace(base.scene, resolution=(64, 64), diskInner=3.0, diskOuter=30.0),
                            camera=dataclasses.replace(base.camera, position=(0.0, 0.0, -20.0), lookAt=(0.0, 0.0, 0.0), upVec=(0.0, 1.0, 0.0)))
    elif case == "cam_on_y_axis":
        cfg = config.Config(scene=dataclasses.replace(base.scene, resolution=(33, 33)),
                            camera=dataclasses.replace(base.camera, position=(0.0, 5.0, 0.0), lookAt=(0.0, 0.0, 0.0), upVec=(0.0, 0.0, 1.0)))
    else:
        cfg = config.Config(scene=dataclasses.replace(base.scene, resolution=(32, 32), stepSize=0.1),
                            camera=dataclasses.replace(base.camera, position=(3.0, 0.5, 0.0), lookAt=(0.0, 0.0, 3.0)))
    rnd.set_stars(stars40k)
    rnd.set_option("trace_variant", 0)
    img = rnd.render(cfg)
    ref, rsteps = po.render(cfg, po.Tree(stars40k))
    assert np.abs(rgb(img) - ref).max() < TOL
    assert rnd.last_stats["steps"] == rsteps - img.shape[0] * img.shape[1]


def test_full_size_rows_default_1080p(rnd, scenes_dir):
    # BASELINE config 2: default.yaml 1920x1080, no star map.  Oracle on a band of rows.
    cfg = config.load_config(f"{scenes_dir}/default.yaml")
    rnd.set_stars(None)
    whole = rnd.render(cfg)
    assert whole.shape == (1080, 1920, 4) and np.isfinite(whole).all()
    for r0 in (0, 537, 1076):
        ref, _ = po.render(cfg, None, r0, r0 + 4)
        assert np.abs(rgb(whole[r0:r0 + 4]) - ref).max() < TOL
    # idempotence: same call, same bits
    np.testing.assert_array_equal(rnd.render(cfg), whole)


def test_full_size_rows_default_aa_4096(rnd, scenes_dir):
    # BASELINE config 3 (headline): default-aa.yaml at 4096x4096 with x4 supersampling + stars + bloom
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default-aa.yaml"), 4096, 4096)
    stars = starmap.synthetic_stars()  # N = 468 861, seed 20190412
    rnd.set_stars(stars)
    tree = po.Tree(stars)
    for variant in (0, 1):
        rnd.set_option("trace_variant", variant)
        for r0 in (1000, 2046):
            band = rnd.render(cfg, r0, r0 + 2)
            ref, _ = po.render(cfg, tree, r0, r0 + 2)
            assert np.abs(rgb(band) - ref).max() < TOL
    rnd.set_option("trace_variant", 0)
    full = rnd.do_render(cfg)   # render + bloom at full size (the whole frame is checked against the
    st = rnd.last_stats         # oracle in test_bloom_full_frame_4096_vs_oracle)
    assert st["rays"] == 4 * 4096 * 4096 and st["capped"] == 0
    assert np.isfinite(full).all()


def test_whole_headline_frame_vs_oracle(rnd, scenes_dir):
    """EVERY pixel of the 4096x4096 default-aa frame (67 M rays, stars, disk, photon ring) against the
    oracle, which takes about a minute on the box's host cores.  The contract is 1e-4 per channel; the
    count of pixels above tighter bars is printed so that a drift shows up long before it matters."""
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default-aa.yaml"), 4096, 4096)
    stars = starmap.synthetic_stars()
    rnd.set_stars(stars)
    rnd.set_option("trace_variant", 6)
    got = rnd.render(cfg)
    st = rnd.last_stats
    ref, rsteps = po.render(cfg, po.Tree(stars))
    err = np.abs(rgb(got) - ref).max(axis=2)
    print(f"whole frame: max err {err.max():.3e}; pixels > 1e-4: {(err > 1e-4).sum()}, > 1e-5: {(err > 1e-5).sum()}, "
          f"> 1e-6: {(err > 1e-6).sum()} of {err.size}")
    assert err.max() < TOL
    assert (err > 1e-5).sum() <= 16
    assert st["steps"] == rsteps - 4 * 4096 * 4096     # identical step counts (minus the unused last step of every ray)


def test_full_size_rows_lensing_disk_8192(rnd, scenes_dir):
    # BASELINE config 4: lensing-disk.yaml at 8192x8192 (x4 supersampling): bands of rows vs the oracle,
    # and the bloom of an 8192-wide frame (the 16-pixels-per-thread variant of the bloom kernel)
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/lensing-disk.yaml"), 8192, 8192)
    stars = starmap.synthetic_stars()
    rnd.set_stars(stars)
    rnd.set_option("trace_variant", 4)
    tree = po.Tree(stars)
    for r0 in (4095, 6000):
        band = rnd.render(cfg, r0, r0 + 1)
        ref, _ = po.render(cfg, tree, r0, r0 + 1)
        assert np.abs(rgb(band) - ref).max() < TOL
    # bloom at this width: r = 8192 // 25 = 327; compare a 8192 x 40 strip against the oracle
    rng = np.random.default_rng(11)
    img = np.ones((40, 8192, 4), dtype=np.float32)
    img[..., :3] = rng.uniform(0, 1, (40, 8192, 3)).astype(np.float32)
    got = rnd.bloom(0.15, 25, img)
    ref = po.bloom(0.15, 25, rgb(img))
    assert np.abs(rgb(got) - ref).max() < 1e-5


def test_animation_frame_1080p_supersampled(rnd, scenes_dir):
    # BASELINE config 5: a frame of animations/default-ani.yaml (1920x1080, supersampling, bloom 0.7)
    from blackstar_b200 import animation
    a = animation.load_animation(os.path.join(os.path.dirname(scenes_dir), "animations", "default-ani.yaml"))
    cfg = animation.generate_frames(a)[200]
    stars = starmap.synthetic_stars(100000, seed=6)
    rnd.set_stars(stars)
    img = rnd.render(cfg)
    assert img.shape == (1080, 1920, 4)
    ref, _ = po.render(cfg, po.Tree(stars), 500, 503)
    assert np.abs(rgb(img[500:503]) - ref).max() < TOL
    full = rnd.do_render(cfg)
    assert np.isfinite(full).all() and (rgb(full) >= rgb(img) - 1e-6).all()


@pytest.mark.parametrize("seed", [3, 19, 40, 41, 57, 77, 90])
def test_random_scenes_match_oracle(rnd, stars40k, seed):
    # the randomized scenes of tests/test_hostcheck.py (cameras in the disk plane, on the polar axis,
    # looking straight at the hole ...) through the real kernels; seed 19 contains an exactly radial ray
    import importlib.util
    spec = importlib.util.spec_from_file_location("th", os.path.join(HERE, "test_hostcheck.py"))
    th = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(th)
    cfg = th._random_config(np.random.default_rng(1000 + seed))
    rnd.set_stars(stars40k)
    ref, rsteps = po.render(cfg, po.Tree(stars40k))
    nrays = ref.shape[0] * ref.shape[1] * (4 if cfg.scene.supersampling else 1)
    for variant in (6, 1):
        rnd.set_option("trace_variant", variant)
        img = rnd.render(cfg)
        assert np.abs(rgb(img) - ref).max() < TOL
        assert rnd.last_stats["steps"] == rsteps - nrays
    rnd.set_option("trace_variant", 6)


@pytest.mark.parametrize("scene", SCENES)
def test_every_scene_at_native_resolution(rnd, scenes_dir, scene):
    # every scene of scenes/ at the resolution it ships with (supersampling as configured), full
    # 468 861-star catalogue: two bands of rows against the oracle, bloom sanity on the whole frame
    cfg = config.load_config(f"{scenes_dir}/{scene}.yaml")
    W, H = cfg.scene.resolution
    stars = starmap.synthetic_stars()
    rnd.set_stars(stars)
    rnd.set_option("trace_variant", 6)
    img = rnd.render(cfg)
    assert img.shape == (H, W, 4) and np.isfinite(img).all() and rnd.last_stats["capped"] == 0
    tree = _full_tree(stars)
    rng = np.random.default_rng(len(scene))
    for r0 in (int(rng.integers(0, H - 2)), H // 2):
        ref, _ = po.render(cfg, tree, r0, r0 + 2)
        assert np.abs(rgb(img[r0:r0 + 2]) - ref).max() < TOL
    full = rnd.do_render(cfg)
    assert np.isfinite(full).all() and (rgb(full) >= rgb(img) - 1e-6).all()


_TREE_CACHE = {}


def _full_tree(stars):
    if "t" not in _TREE_CACHE:
        _TREE_CACHE["t"] = po.Tree(stars)
    return _TREE_CACHE["t"]
