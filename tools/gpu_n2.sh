#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (2 GPUs visible)"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== torchrun bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n2.err | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<PY
import json
def load(p):
    for l in open(p):
        if l.startswith('{'): return json.loads(l)
base=None
for n in (1,2):
    d=load('gpurun_out/bench_n%d.json'%n)
    if base is None: base=d
    print(n, 'value %.1f ms %.3f x%.2f | e2e %.1f x%.2f | pipelined %.1f | launches %d | trace(rank0) %.2f ms | fp64 peak %.2f | tiles %s' % (d['value'], d['ms_per_step'], d['value']/base['value'], d['e2e']['value'], d['e2e']['value']/base['e2e']['value'], d['e2e']['pipelined']['value'], d['gpu_launches'], d['roofline']['launch_ms'], d['roofline_fp64']['peak'], d['config'].get('tiles')))
PY
head -c 300 gpurun_out/bench_n2.json
