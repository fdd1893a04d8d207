#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 | tee gpurun_out/pytest_gpu_n2.txt
