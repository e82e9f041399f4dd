#!/bin/bash
# Same-box A/B: previous commit's library vs the current one (features on / off) vs the current one without the epilogue's split stores.
mkdir -p gpurun_out
run() {  # label lib xb xprod batch
  FLOWSE_XB=$3 FLOWSE_XPROD=$4 FLOWSE_LIB=$PWD/flowmse_b200/$2 timeout 600 python bench.py --steps 10 --batch $5 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab2.json 2> gpurun_out/ab2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab2.json"))
print("$1 B=$5: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"])
PY
}
for rep in 1 2 3; do
  run "old            " libflowse_old.so 0 0 1
  run "new xb0 xp0    " libflowse.so 0 0 1
  run "new xb1 xp0    " libflowse.so 1 0 1
  run "new xb1 xp1    " libflowse.so 1 1 1
  run "nox xb1 xp0    " libflowse_nox.so 1 0 1
done
run "old            " libflowse_old.so 0 0 8
run "new xb1 xp0    " libflowse.so 1 0 8
run "new xb1 xp1    " libflowse.so 1 1 8
run "nox xb1 xp0    " libflowse_nox.so 1 0 8
