#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_b.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}); print(d["e2e"])
PY
tail -3 gpurun_out/bench_b.err
