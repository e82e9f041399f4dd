#!/bin/bash
# Same-box A/B of two builds of the library: flowmse_b200/libflowse_old.so (an earlier commit) vs the current one.
mkdir -p gpurun_out
for rep in 1 2 3; do
for lib in libflowse_old.so libflowse.so; do
  FLOWSE_LIB=$PWD/flowmse_b200/$lib timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab_$lib.$rep.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$lib.$rep.json"))
print("$lib rep $rep: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"], "frac", round(d["roofline"]["frac"],4))
PY
done
done
for lib in libflowse_old.so libflowse.so; do
  FLOWSE_LIB=$PWD/flowmse_b200/$lib timeout 600 python bench.py --steps 10 --batch 8 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab_b8_$lib.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_b8_$lib.json"))
print("B=8 $lib: value",round(d["value"]),"ms",round(d["ms_per_step"],3))
PY
done
