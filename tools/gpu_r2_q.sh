#!/bin/bash
# Round 2, call Q: forked graph branches A/B + the whole GPU suite.
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.json
for f in tests/test_gpu_*.py; do
  name=$(basename $f .py)
  timeout 1500 python -m pytest $f -m gpu -q -s > gpurun_out/q_$name.log 2>&1
  echo "== $f: exit $?"; tail -2 gpurun_out/q_$name.log | cut -c1-200
done
for fk in 0 1; do
  FLOWSE_FORK=$fk timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/q_bench_fork$fk.json 2> gpurun_out/q_bench_fork$fk.err
  python - <<PY
import json
d=json.load(open("gpurun_out/q_bench_fork$fk.json"))
print("fork $fk: value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"])
PY
done
