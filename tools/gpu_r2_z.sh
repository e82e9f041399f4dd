#!/bin/bash
# Round 2, call Z: shortcut tiles through the weight ring (FLOWSE_XB) x producers write the shortcut operand (FLOWSE_XPROD).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_forward.py tests/test_gpu_ops.py -m gpu -q -k "fused_operand or overflow or golden or euler or conv" > gpurun_out/z_parity.log 2>&1
echo "== parity exit $?"; tail -3 gpurun_out/z_parity.log | cut -c1-300
for rep in 1 2; do
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  FLOWSE_XB=$1 FLOWSE_XPROD=$2 timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/z_bench_xb$1_xp$2_$rep.json 2> gpurun_out/z_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/z_bench_xb$1_xp$2_$rep.json"))
print("xb $1 xprod $2 rep $rep: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"], "frac", round(d["roofline"]["frac"],4))
PY
done
done
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  FLOWSE_XB=$1 FLOWSE_XPROD=$2 timeout 600 python bench.py --steps 10 --batch 4 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/z_bench_b4_xb$1_xp$2.json 2> gpurun_out/z_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/z_bench_b4_xb$1_xp$2.json"))
print("B=4 xb $1 xprod $2: value",round(d["value"]),"ms",round(d["ms_per_step"],3))
PY
done
FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2>&1 | grep "halo dbg" | cut -c1-330 | sed -n '30,40p'
