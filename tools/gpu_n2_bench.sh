#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n2.err | tail -5
head -c 200 gpurun_out/bench_n2.json; echo
python - <<PY
import json
for l in open('gpurun_out/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['config']['tiles'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 2 --warmup 1 --impl reference --res 512 512 2>/dev/null | cut -c1-200
