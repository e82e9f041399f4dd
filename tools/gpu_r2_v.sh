#!/bin/bash
# Round 2, call V: plain preparation as a prologue phase of the per-tap conv kernel (fuse_prep = 2): parity, A/B, phases.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "fused_operand" > gpurun_out/v_parity.log 2>&1
echo "== fused parity exit $?"; tail -4 gpurun_out/v_parity.log | cut -c1-300
for fp in 1 2; do
  FLOWSE_FUSE_PREP=$fp timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/v_bench_fp$fp.json 2> gpurun_out/v_bench_fp$fp.err
  echo "bench fp=$fp exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/v_bench_fp$fp.json"))
print("fuse_prep $fp: value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"], d["roofline"].get("nfe_ms_by_kernel_family"))
PY
done
FLOWSE_FUSE_PREP=2 timeout 600 python bench.py --steps 10 --batch 4 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/v_bench_b4.json 2> gpurun_out/v_bench_b4.err
python - <<PY
import json
d=json.load(open("gpurun_out/v_bench_b4.json"))
print("fuse_prep 2 B=4: value",round(d["value"]),"ms",round(d["ms_per_step"],3))
PY
FLOWSE_FUSE_PREP=2 FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2>&1 | grep "conv dbg" | sort | uniq -c | sort -rn | head -12
