#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitizer_workload.py 2>&1 | tail -5 | tee gpurun_out/sanitizer_$tool.txt
done
