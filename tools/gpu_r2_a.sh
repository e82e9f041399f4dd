#!/bin/bash
# Round 2, call A: whole GPU suite (one process per file), default bench line, reference arm (short).
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.json
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_*.py; do
  name=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q -s > gpurun_out/$name.log 2>&1
  echo "== $f: exit $?"; tail -4 gpurun_out/$name.log
done
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench exit $?"; head -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
