#!/bin/bash
# Which thread's clock reads matter: base (both), p0 (producer without), m0 (MMA issuer without).
mkdir -p gpurun_out
run() {  # label lib batch
  FLOWSE_LIB=$PWD/flowmse_b200/$2 timeout 600 python bench.py --steps 10 --batch $3 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab5.json 2> gpurun_out/ab5.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab5.json"))
print("$1 B=$3: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"])
PY
}
for rep in 1 2 3; do
  run "base" libflowse.so 1
  run "p0  " libflowse_p0.so 1
  run "m0  " libflowse_m0.so 1
done
