#!/bin/bash
# Sensitivity of the halo kernel to the depth of its weight ring: 3 stages (current) vs 2.
mkdir -p gpurun_out
run() {  # label lib batch
  FLOWSE_LIB=$PWD/flowmse_b200/$2 timeout 600 python bench.py --steps 10 --batch $3 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab8.json 2> gpurun_out/ab8.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab8.json"))
print("$1 B=$3: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"], "frac", round(d["roofline"]["frac"],4))
PY
}
for rep in 1 2; do
  run "B ring 3" libflowse.so 1
  run "B ring 2" libflowse_b2.so 1
done
FLOWSE_LIB=$PWD/flowmse_b200/libflowse_b2.so FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2>&1 | grep "halo dbg" | cut -c1-330 | sed -n '1,2p;32,36p'
