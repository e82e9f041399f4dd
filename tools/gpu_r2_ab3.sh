#!/bin/bash
# Same-box A/B: previous commit's library vs the current one with shortcut tiles through the weight ring off / on.
mkdir -p gpurun_out
run() {  # label lib xb batch
  FLOWSE_XB=$3 FLOWSE_LIB=$PWD/flowmse_b200/$2 timeout 600 python bench.py --steps 10 --batch $4 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab3.json 2> gpurun_out/ab3.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab3.json"))
print("$1 B=$4: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"])
PY
}
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_forward.py -m gpu -q -k "fused_operand or golden or euler" > gpurun_out/ab3_parity.log 2>&1
echo "== parity exit $?"; tail -2 gpurun_out/ab3_parity.log | cut -c1-200
for rep in 1 2 3; do
  run "old    " libflowse_old.so 0 1
  run "new xb0" libflowse.so 0 1
  run "new xb1" libflowse.so 1 1
done
run "old    " libflowse_old.so 0 8
run "new xb0" libflowse.so 0 8
run "new xb1" libflowse.so 1 8
