#!/bin/bash
# Run each GPU test file in its own process (a trapped kernel poisons the CUDA context) and keep the logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/$name.log
  echo "== $f: exit ${PIPESTATUS[0]}"; tail -25 gpurun_out/$name.log
done
