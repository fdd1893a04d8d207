"""Time experimental builds of the library (blackstar_b200/csrc/exp/libexp_*.so) on the headline frame."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, numpy as np
sys.path.insert(0, %r)
import torch
from blackstar_b200 import config, starmap
from blackstar_b200.render import Renderer
base = config.with_resolution(config.load_config(os.path.join(%r, "scenes/default-aa.yaml")), 4096, 4096)
stars = starmap.synthetic_stars()
buf = torch.empty((4096, 4096, 4), dtype=torch.float32, device="cuda")
with Renderer(devices=[0]) as r:
    r.set_stars(stars)
    for _ in range(2): r.render_device(base, buf.data_ptr(), want_stats=True)
    ts = [r.render_device(base, buf.data_ptr(), want_stats=True) for _ in range(4)]
    ms = min(t["trace_ms"] for t in ts)
    band = buf[2040:2056].cpu().numpy().astype(np.float64)
    print("%%.3f ms  steps %%d  checksum %%.12f" %% (ms, ts[0]["steps"], band.sum()))
''' % (ROOT, ROOT)
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "blackstar_b200", "csrc", "exp", "libexp_*.so"))) + [None]
for lib in libs:
    env = dict(os.environ)
    if lib: env["BLACKSTAR_B200_LIB"] = lib
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib) if lib else 'default':28s} {out.stdout.strip()} {out.stderr.strip()[-200:]}")
