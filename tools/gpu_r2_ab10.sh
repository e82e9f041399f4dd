#!/bin/bash
# Same-box A/B/C: rows per straight-line batch of the transform warps (XF_NB = 6 current, 3, 2) with 128 registers per thread.
mkdir -p gpurun_out
run() {  # label lib batch
  FLOWSE_LIB=$PWD/flowmse_b200/$2 timeout 600 python bench.py --steps 10 --batch $3 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab10.json 2> gpurun_out/ab10.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab10.json"))
print("$1 B=$3: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"])
PY
}
for rep in 1 2 3; do
  run "NB=6" libflowse.so 1
  run "NB=3" libflowse_nb3.so 1
  run "NB=2" libflowse_nb2.so 1
done
run "NB=6" libflowse.so 8
run "NB=3" libflowse_nb3.so 8
run "NB=2" libflowse_nb2.so 8
