#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "== rk4 probe"; ./tools/rk4_pipe_probe | tee gpurun_out/rk4_pipe_probe2.txt
echo "== trace experiments"
timeout 600 python tools/trace_experiments.py | tee gpurun_out/trace_experiments3.txt
for v in 6 4; do
echo "== bench variant $v"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant $v > gpurun_out/bench7_v$v.json 2> gpurun_out/bench7_v$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench7_v$v.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["e2e"]["pipelined"]["value"], d["roofline"]["launch_ms"])
PY
done
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_tiles" -s 1 -c 1 -o gpurun_out/prof_r01f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --variant 6 > gpurun_out/ncu_full_bench7.log 2>&1
ls -la gpurun_out/prof_r01f.ncu-rep
