#!/bin/bash
# Same-box A/B: library built with an earlier version of one source file (pre) vs the current one.
mkdir -p gpurun_out
run() {  # label lib batch
  FLOWSE_LIB=$PWD/flowmse_b200/$2 timeout 600 python bench.py --steps 10 --batch $3 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab9.json 2> gpurun_out/ab9.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab9.json"))
print("$1 B=$3: value",round(d["value"]),"ms",round(d["ms_per_step"],3), d["roofline"]["nfe_ms_by_kernel_family"])
PY
}
for rep in 1 2 3; do
  run "pre    " libflowse_pre.so 1
  run "current" libflowse.so 1
done
run "pre    " libflowse_pre.so 8
run "current" libflowse.so 8
