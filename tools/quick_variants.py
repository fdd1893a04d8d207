import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blackstar_b200 import config, starmap
from blackstar_b200.render import Renderer
base = config.with_resolution(config.load_config("scenes/default-aa.yaml"), 4096, 4096)
stars = starmap.synthetic_stars()
buf = torch.empty((4096, 4096, 4), dtype=torch.float32, device="cuda")
with Renderer(devices=[0]) as r:
    r.set_stars(stars)
    for v in (6, 0, 4, 6, 0, 4):
        r.set_option("trace_variant", v)
        for _ in range(2): r.render_device(base, buf.data_ptr(), want_stats=True)
        ms = min(r.render_device(base, buf.data_ptr(), want_stats=True)["trace_ms"] for _ in range(4))
        print(f"variant {v}: {ms:.3f} ms")
