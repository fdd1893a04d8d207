"""Multi-GPU inside ONE process (bsb_render_full: row tiles on every GPU of the ctx, one NCCL
gather on GPU 0, bloom on GPU 0).  Needs >= 2 devices; on a 1-GPU box only the 1-GPU
identity is checked."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from blackstar_b200 import config, starmap  # noqa: E402
from blackstar_b200.render import Renderer  # noqa: E402


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("scene,res", [("default-aa", (256, 145)), ("lensing-disk", (200, 126))])
def test_render_full_is_identical_on_1_and_n_gpus(scenes_dir, scene, res):
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
    for k in sorted({2, n}):
        with Renderer(n_gpus=k) as rk:
            rk.set_stars(stars)
            got = rk.do_render(cfg)          # equal row tiles
            st = rk.last_stats
            got8 = rk.do_render_srgb8(cfg)   # tiles re-cut from the measured per-GPU rates
            again = rk.do_render(cfg)
        assert st["n_gpus"] == k and st["launches"] == 2 * k + 2
        np.testing.assert_array_equal(got, ref)
        np.testing.assert_array_equal(got8, ref8)
        np.testing.assert_array_equal(again, ref)
