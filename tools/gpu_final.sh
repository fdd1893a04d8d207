#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_final.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}); print(d["e2e"]); print(d["roofline"]); print(d["roofline_fp64"]["frac"], d["roofline_fp64"]["executed"]["pipe_frac"]); print(d["roofline_bloom"]["launch_ms"], d["roofline_bloom"]["frac"]); print(d["cpu_baseline"])
PY
tail -3 gpurun_out/bench_final.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
grep -c trace gpurun_out/launches_final.csv
