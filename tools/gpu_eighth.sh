#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== trace experiments"
timeout 600 python tools/trace_experiments.py | tee gpurun_out/trace_experiments4.txt
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench8.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}); print(d["e2e"]); print(d["roofline"]["launch_ms"], d["roofline_fp64"]["executed"]["pipe_frac"]); print(d["roofline_bloom"]); print(d["cpu_baseline"])
PY
tail -3 gpurun_out/bench8.err
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_tiles|box3" -s 3 -c 3 -o gpurun_out/prof_r01g \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench8.log 2>&1
ls -la gpurun_out/prof_r01g.ncu-rep
