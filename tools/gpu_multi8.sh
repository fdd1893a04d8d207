#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== in-library multi-GPU test"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for n in $N 4; do
echo "== torchrun bench N=$n"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n$n.err | tail -5
done
echo "== bench N=1 (same box)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<PY
import json
def load(p):
    for l in open(p):
        if l.startswith('{'): return json.loads(l)
base=None
for n in (1,4,$N):
    d=load('gpurun_out/bench_n%d.json'%n)
    if base is None: base=d
    print(n, 'value %.1f ms %.3f x%.2f | e2e %.1f x%.2f | pipelined %.1f | launches %d | trace(rank0) %.2f ms | tiles %s' % (d['value'], d['ms_per_step'], d['value']/base['value'], d['e2e']['value'], d['e2e']['value']/base['e2e']['value'], d['e2e']['pipelined']['value'], d['gpu_launches'], d['roofline']['launch_ms'], d['config'].get('tiles')))
PY
