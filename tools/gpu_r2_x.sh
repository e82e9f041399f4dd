#!/bin/bash
# Round 2, call X: how the transform warps load the fp32 activations (L1-allocating vs L2-only vs no-allocate vs streaming).
mkdir -p gpurun_out
for rep in 1 2; do
for m in 0 1 2 3; do
  lib=flowmse_b200/libflowse.so; [ $m != 0 ] && lib=flowmse_b200/libflowse_ld$m.so
  FLOWSE_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/x_bench_ld${m}_$rep.json 2> gpurun_out/x_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/x_bench_ld${m}_$rep.json"))
print("ldmode $m rep $rep: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"], "frac", round(d["roofline"]["frac"],4))
PY
done
done
for m in 0 1 2; do
  lib=flowmse_b200/libflowse.so; [ $m != 0 ] && lib=flowmse_b200/libflowse_ld$m.so
  FLOWSE_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --batch 4 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/x_bench_b4_ld$m.json 2> gpurun_out/x_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/x_bench_b4_ld$m.json"))
print("B=4 ldmode $m: value",round(d["value"]),"ms",round(d["ms_per_step"],3))
PY
done
FLOWSE_LIB=$PWD/flowmse_b200/libflowse_ld1.so FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2>&1 | grep "halo dbg" | cut -c1-330 | sed -n '1,2p;30,40p'
