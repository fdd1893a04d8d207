#!/bin/bash
mkdir -p gpurun_out
./tools/fp64_pipe_probe | tee gpurun_out/fp64_pipe_probe.txt
echo "== compute-sanitizer memcheck (small frames, all schedules)"
cat > san_tmp.py <<PY
import numpy as np
from blackstar_b200 import config, starmap
from blackstar_b200.render import Renderer
cfg = config.with_resolution(config.load_config("scenes/default-aa.yaml"), 61, 35)
with Renderer(devices=[0]) as r:
    r.set_stars(starmap.synthetic_stars(20000, seed=3))
    for v in (0, 1, 2, 3, 4, 6):
        r.set_option("trace_variant", v)
        img = r.do_render(cfg)
    u8 = r.do_render_srgb8(cfg)
    cfg2 = config.with_resolution(config.load_config("scenes/default.yaml"), 77, 41)
    img2 = r.do_render(cfg2)
print("sanitizer workload done", float(img.mean()), float(img2.mean()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python san_tmp.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python san_tmp.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck.txt
rm -f san_tmp.py
