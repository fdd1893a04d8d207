"""Small workload for compute-sanitizer: every trace variant, supersampling on/off, star lookups,
bloom at several line lengths (1, 2, 4, 8 pixels per thread; odd sizes), sRGB8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from blackstar_b200 import config, starmap
from blackstar_b200.render import Renderer

cfg = config.with_resolution(config.load_config("scenes/default-aa.yaml"), 61, 35)
cfg2 = config.with_resolution(config.load_config("scenes/default.yaml"), 77, 41)
with Renderer(devices=[0]) as r:
    r.set_stars(starmap.synthetic_stars(150000, seed=3))   # depth 15: top levels + one record group
    for v in (0, 1, 2, 3, 4, 6):
        r.set_option("trace_variant", v)
        img = r.do_render(cfg)
    u8 = r.do_render_srgb8(cfg)
    img2 = r.do_render(cfg2)
    rng = np.random.default_rng(0)
    for (h, w, div) in ((5, 300, 25), (33, 600, 25), (17, 1100, 25), (9, 2300, 25), (3, 4100, 25), (64, 40, 7), (37, 53, 5)):
        a = np.ones((h, w, 4), dtype=np.float32)
        a[..., :3] = rng.uniform(0, 1, (h, w, 3)).astype(np.float32)
        b = r.bloom(0.3, div, a)
print("sanitizer workload done", float(img.mean()), float(img2.mean()), float(b.mean()))
