"""Multi-GPU: bsb_render_full inside ONE process (row tiles on every GPU of the ctx, horizontal bloom
on the tiles, one NCCL all-to-all into column bands, vertical bloom on the bands, N parallel copies
into the host frame) and the same pipeline with one process per GPU (blackstar_b200/dist.py under
torchrun).  Needs >= 2 devices: on a 1-GPU box only the 1-GPU half runs and the rest is skipped
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from blackstar_b200 import config, starmap  # noqa: E402
from blackstar_b200.render import Renderer  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("scene,res", [("default-aa", (256, 146)), ("lensing-disk", (200, 126)), ("default", (1920, 1080))])
def test_render_full_on_1_and_n_gpus(scenes_dir, scene, res):
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/{scene}.yaml"), *res)
    stars = starmap.synthetic_stars(50000, seed=8)
    with Renderer(devices=[0]) as r1:
        r1.set_stars(stars)
        ref = r1.do_render(cfg)
        pre = r1.render(cfg)
        ref8 = r1.do_render_srgb8(cfg)
    assert np.isfinite(ref).all()
    assert not np.array_equal(ref, pre)  # bloom did something
    n = _ngpu()
    if n < 2:
        pytest.skip("one GPU visible: the N-GPU half needs gpurun --gpus 2")
    import torch
    nobloom = config.Config(scene=__import__("dataclasses").replace(cfg.scene, bloomStrength=0.0), camera=cfg.camera)
    for k in sorted({2, min(n, 3), n}):
        with Renderer(n_gpus=k) as rk:
            rk.set_stars(stars)
            got = rk.do_render(cfg)          # equal row tiles
            st = rk.last_stats
            got8 = rk.do_render_srgb8(cfg)   # tiles re-cut from the measured per-GPU rates
            pinned = torch.empty(got.shape, dtype=torch.float32, pin_memory=True)
            again = rk.do_render(cfg, out=pinned.numpy()).copy()
            raw = rk.do_render(nobloom)
        assert st["n_gpus"] == k and st["launches"] == 5 * k
        # the trace is bit-exact whatever the tiling; the bloom's prefix sums are chunked differently
        np.testing.assert_array_equal(raw, pre)
        assert np.abs(got - ref).max() < 2e-6
        assert np.abs(again - ref).max() < 2e-6
        d8 = np.abs(got8.astype(int) - ref8.astype(int))
        assert d8.max() <= 1 and (d8 != 0).mean() < 1e-4


def test_sides_above_8192_on_n_gpus(scenes_dir):
    """Bloom lines longer than the shared-memory kernel takes: on a multi-GPU ctx the tiles are traced everywhere,
    copied to the first GPU over NVLink and bloomed there by the long-line path.  Against the 1-GPU render."""
    cfg = config.with_resolution(config.load_config(f"{scenes_dir}/default.yaml"), 8208, 24)
    stars = starmap.synthetic_stars(50000, seed=8)
    with Renderer(devices=[0]) as r1:
        r1.set_stars(stars)
        ref = r1.do_render(cfg)
        ref8 = r1.do_render_srgb8(cfg)
    n = _ngpu()
    if n < 2:
        pytest.skip("one GPU visible: needs gpurun --gpus 2")
    with Renderer(n_gpus=min(n, 3)) as rk:
        rk.set_stars(stars)
        got = rk.do_render(cfg)
        got8 = rk.do_render_srgb8(cfg)
    np.testing.assert_array_equal(got, ref)        # same kernels on the same frame: bit-identical
    np.testing.assert_array_equal(got8, ref8)


@pytest.mark.parametrize("res,bloom", [((640, 362), True), ((1920, 1080), True), ((300, 168), False)])
def test_one_process_per_gpu_pipeline(scenes_dir, tmp_path, res, bloom):
    """blackstar_b200.dist.DistributedFrame under torchrun on every visible GPU (>= 2): the frame that lands
    in the shared host buffer against the 1-GPU render of the same scene."""
    n = _ngpu()
    if n < 2:
        pytest.skip("one GPU visible: needs gpurun --gpus 2")
    out = str(tmp_path / "dist.json")
    for world in sorted({2, n}):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + world), os.path.join(ROOT, "tools", "dist_check.py"), "--res", str(res[0]), str(res[1]),
               "--out", out] + ([] if bloom else ["--no-bloom"])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rep = json.load(open(out))
        assert rep["world"] == world
        assert rep["max_abs_err_f32"] < 2e-6, rep
        assert rep["srgb8_max_diff"] <= 1 and rep["srgb8_frac_diff"] < 1e-4, rep
