"""Where does the trace kernel lose FP64-pipe time?  G steps/s for variants of the headline frame
(the bare RK4 stream sustains ~272 G steps/s, tools/rk4_pipe_probe.cu)."""
import dataclasses, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blackstar_b200 import config, starmap
from blackstar_b200.render import Renderer

base = config.with_resolution(config.load_config("scenes/default-aa.yaml"), 4096, 4096)
stars = starmap.synthetic_stars()
buf = torch.empty((4096, 4096, 4), dtype=torch.float32, device="cuda")
variants = {
    "headline (ss, stars, disk)": (base, stars),
    "no stars": (base, None),
    "no disk": (config.Config(scene=dataclasses.replace(base.scene, diskOpacity=0.0), camera=base.camera), stars),
    "no stars, no disk": (config.Config(scene=dataclasses.replace(base.scene, diskOpacity=0.0), camera=base.camera), None),
    "no ss (8192^2 grid)": (config.Config(scene=dataclasses.replace(base.scene, supersampling=False, resolution=(4096, 4096)), camera=base.camera), stars),
    "fartheraway cam (468 steps/ray)": (config.with_resolution(config.load_config("scenes/fartheraway.yaml"), 2048, 2048), stars),
    "closeup cam (16% captured)": (config.with_resolution(config.load_config("scenes/closeup.yaml"), 4096, 3072), stars),
}
with Renderer(devices=[0]) as r:
    for v in (6, 1, 3):
        r.set_option("trace_variant", v)
        for name, (cfg, st) in variants.items():
            r.set_stars(st)
            W, H = cfg.scene.resolution
            for _ in range(2):
                s = r.render_device(cfg, buf.data_ptr(), want_stats=True)
            ms = min(r.render_device(cfg, buf.data_ptr(), want_stats=True)["trace_ms"] for _ in range(3))
            print(f"variant {v} | {name:34s} {ms:8.3f} ms  {s['rays']/ms/1e3:8.1f} Mrays/s  {s['steps']/s['rays']:6.1f} steps/ray  {s['steps']/ms/1e6:7.1f} Gsteps/s")
