#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
for v in 4 5 0; do
  echo "== bench variant $v"
  timeout 600 python bench.py --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench3_v$v.json 2> gpurun_out/bench3_v$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench3_v$v.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], "trace_ms", d["roofline"]["launch_ms"], "bloom_ms", d.get("roofline_bloom",{}).get("launch_ms"))
PY
  tail -3 gpurun_out/bench3_v$v.err
done
echo "== ncu (trace v5 + bloom)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace|box3" -s 3 -c 3 -o gpurun_out/prof_r01d \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --variant 5 > gpurun_out/ncu_full_bench3.log 2>&1
ls -la gpurun_out | tail -4
