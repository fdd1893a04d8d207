#!/usr/bin/env python
"""One render of the headline frame (default-aa at WxH, full star catalogue) for ncu captures of the trace
kernel; prints the RK4 step count of the launch.  usage: python tools/trace_workload.py [W H] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from blackstar_b200 import config, starmap  # noqa: E402
from blackstar_b200.render import Renderer  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", "default-aa.yaml")), W, H)
buf = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
with Renderer(devices=[0]) as r:
    r.set_stars(starmap.synthetic_stars())
    for _ in range(reps):
        st = r.render_device(cfg, buf.data_ptr(), want_stats=True)
print(json.dumps({"W": W, "H": H, "rays": st["rays"], "steps": st["steps"], "star_hits": st["star_hits"], "trace_ms": st["trace_ms"]}))
