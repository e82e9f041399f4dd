#!/bin/bash
# Same-box A/B/C: 640-thread fused halo kernel (old) vs 512 threads (epw4) vs 512 threads + two register sets (new).
mkdir -p gpurun_out
run() {  # label lib batch
  FLOWSE_LIB=$PWD/flowmse_b200/$2 timeout 600 python bench.py --steps 10 --batch $3 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab7.json 2> gpurun_out/ab7.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab7.json"))
print("$1 B=$3: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"], "frac", round(d["roofline"]["frac"],4))
PY
}
for rep in 1 2 3 4; do
  run "old " libflowse_old.so 1
  run "epw4" libflowse_epw4.so 1
  run "new " libflowse.so 1
done
for rep in 1 2; do
  run "old " libflowse_old.so 8
  run "epw4" libflowse_epw4.so 8
  run "new " libflowse.so 8
done
