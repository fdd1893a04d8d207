#!/bin/bash
# first GPU pass: smoke, bench per trace schedule, parity tests, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
for v in 0 1 2 3; do
  echo "== bench variant $v"
  extra="--no-cpu-baseline"; [ $v = 0 ] && extra=""
  timeout 600 python bench.py --steps 5 --warmup 3 --variant $v $extra > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  tail -c 2500 gpurun_out/bench_v$v.json; tail -3 gpurun_out/bench_v$v.err
done
echo "== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
echo "== ncu full (trace + bloom)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace|box3" -s 3 -c 3 -o gpurun_out/prof_r01 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
