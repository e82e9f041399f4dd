#!/bin/bash
# Round 2, call Y: shortcut operand written by its producers (xprod) - parity, A/B, wait breakdown.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_forward.py -m gpu -q -s -k "fused_operand or overflow or golden or euler" > gpurun_out/y_parity.log 2>&1
echo "== parity exit $?"; grep "parity_r2\] fuse" gpurun_out/y_parity.log | head -8; tail -2 gpurun_out/y_parity.log | cut -c1-300
for rep in 1 2; do
for xp in 0 1; do
  FLOWSE_XPROD=$xp timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/y_bench_xp${xp}_$rep.json 2> gpurun_out/y_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/y_bench_xp${xp}_$rep.json"))
print("xprod $xp rep $rep: value",round(d["value"]),"ms",round(d["ms_per_step"],3), d["roofline"]["nfe_ms_by_kernel_family"], "frac", round(d["roofline"]["frac"],4))
PY
done
done
for xp in 0 1; do
  FLOWSE_XPROD=$xp timeout 600 python bench.py --steps 10 --batch 4 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/y_bench_b4_xp$xp.json 2> gpurun_out/y_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/y_bench_b4_xp$xp.json"))
print("B=4 xprod $xp: value",round(d["value"]),"ms",round(d["ms_per_step"],3))
PY
done
FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2>&1 | grep "halo dbg" | cut -c1-330 | sed -n '18,40p'
