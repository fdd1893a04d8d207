#!/bin/bash
# run with gpurun --gpus N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
echo "== in-library multi-GPU test"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
echo "== torchrun bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1800 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
echo "== bench N=1 (same box)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json
for n in (1,$N):
    d=json.load(open('gpurun_out/bench_n%d.json'%n)); print(n, d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['roofline']['launch_ms'])
"
echo "== CLI"
timeout 300 python -m blackstar_b200 -f -s synthetic -o gpurun_out/cli_out scenes/default.yaml 2>&1 | tail -6
timeout 300 python -m blackstar_b200 -p -f -s synthetic -o gpurun_out/cli_out scenes 2>&1 | tail -4
ls -la gpurun_out/cli_out | head -14
