#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/pytest_gpu.txt
for v in 0 4; do
  echo "== bench variant $v"
  extra="--no-cpu-baseline"; [ $v = 0 ] && extra=""
  timeout 600 python bench.py --steps 5 --warmup 3 --variant $v $extra > gpurun_out/bench2_v$v.json 2> gpurun_out/bench2_v$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench2_v$v.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}, d["e2e"], {k:d["roofline"][k] for k in ("launch_ms","achieved","frac")}, d["roofline_fp64"]["achieved"], d["roofline_fp64"]["peak"], d.get("roofline_bloom",{}).get("launch_ms"), d.get("cpu_baseline"))
PY
  tail -3 gpurun_out/bench2_v$v.err
done
echo "== ncu full (trace + bloom)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace|box3" -s 3 -c 3 -o gpurun_out/prof_r01c \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench2.log 2>&1
ls -la gpurun_out | tail -8
