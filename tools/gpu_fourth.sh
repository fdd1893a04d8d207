#!/bin/bash
mkdir -p gpurun_out
echo "== bench (default variant) with cpu baseline"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench4.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}); print(d["e2e"]); print(d["roofline"]); print(d["roofline_fp64"]); print(d["cpu_baseline"])
PY
tail -3 gpurun_out/bench4.err
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench4_ref.json 2> gpurun_out/bench4_ref.err; cut -c1-400 gpurun_out/bench4_ref.json; tail -3 gpurun_out/bench4_ref.err
echo "== compute-sanitizer memcheck (small frames, all schedules)"
cat > /tmp/san.py <<PY
import numpy as np
from blackstar_b200 import config, starmap
from blackstar_b200.render import Renderer
cfg = config.with_resolution(config.load_config("scenes/default-aa.yaml"), 61, 35)
with Renderer(devices=[0]) as r:
    r.set_stars(starmap.synthetic_stars(20000, seed=3))
    for v in range(6):
        r.set_option("trace_variant", v)
        img = r.do_render(cfg)
    u8 = r.do_render_srgb8(cfg)
    cfg2 = config.with_resolution(config.load_config("scenes/default.yaml"), 77, 41)
    img2 = r.do_render(cfg2)
print("sanitizer workload done", float(img.mean()), float(img2.mean()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck.txt
