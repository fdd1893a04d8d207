"""Small workload for compute-sanitizer: every trace variant, supersampling on/off, star lookups,
bloom at every (threads, pixels/thread, r mod C) family incl. odd sizes and the long-line path, the fused
sRGB8 epilogue, the two bloom halves with segmented columns (the multi-GPU building blocks), pageable and
pinned host copies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from blackstar_b200 import config, starmap
from blackstar_b200.dist import col_bands, even_row_tiles
from blackstar_b200.render import Renderer

cfg = config.with_resolution(config.load_config("scenes/default-aa.yaml"), 61, 35)
cfg2 = config.with_resolution(config.load_config("scenes/default.yaml"), 77, 41)
with Renderer(devices=[0]) as r:
    r.set_stars(starmap.synthetic_stars(150000, seed=3))   # depth 15: top levels + one record group
    for v in (0, 1, 2, 3, 4, 6):
        r.set_option("trace_variant", v)
        img = r.do_render(cfg)
    r.set_option("trace_variant", 6)
    u8 = r.do_render_srgb8(cfg)
    img2 = r.do_render(cfg2)
    rng = np.random.default_rng(0)
    for (h, w, div) in ((5, 300, 25), (33, 600, 25), (17, 1100, 25), (9, 2300, 25), (3, 4100, 25), (64, 40, 7), (37, 53, 5),
                        (2, 8200, 40), (4100, 3, 1), (12, 511, 3), (6, 2049, 9)):
        a = np.ones((h, w, 4), dtype=np.float32)
        a[..., :3] = rng.uniform(0, 1, (h, w, 3)).astype(np.float32)
        b = r.bloom(0.3, div, a)
    # the distributed halves on one GPU: 3 row tiles -> 3 column bands
    H, W, div = 70, 130, 6
    a = torch.rand((H, W, 4), device="cuda")
    tiles, bands = even_row_tiles(H, 3), col_bands(W, 3)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    mids, imgs = [], []
    for r0, r1 in tiles:
        m, t = torch.empty((W, r1 - r0, 4), device="cuda"), torch.empty((W, r1 - r0, 4), device="cuda")
        r.bloom_h_device(W // div, W, r1 - r0, a[r0:r1].data_ptr(), m.data_ptr(), t.data_ptr())
        mids.append(m); imgs.append(t)
    for c0, c1 in bands:
        pm, pi = [m[c0:c1].contiguous() for m in mids], [t[c0:c1].contiguous() for t in imgs]
        band = torch.empty((H, c1 - c0, 4), device="cuda")
        band8 = torch.empty((H, c1 - c0, 3), device="cuda", dtype=torch.uint8)
        r.bloom_v_device(0.3, W // div, H, c1 - c0, [x.data_ptr() for x in pm], [x.data_ptr() for x in pi],
                         [t1 - t0 for t0, t1 in tiles], band.data_ptr(), band8.data_ptr())
    torch.cuda.synchronize()
    r.set_stream(None)
    big = config.with_resolution(config.load_config("scenes/default.yaml"), 1920, 300)
    pinned = torch.empty((300, 1920, 4), dtype=torch.float32, pin_memory=True)
    p1 = r.do_render(big, out=pinned.numpy())
    p2 = r.do_render(big)
print("sanitizer workload done", float(img.mean()), float(img2.mean()), float(b.mean()), float(band.mean()), bool((p1 == p2).all()))
