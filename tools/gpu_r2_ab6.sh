#!/bin/bash
# Same-box A/B: previous build (640-thread fused halo kernel, 2 epilogue warps per quadrant) vs 512 threads / 128 registers.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_forward.py tests/test_gpu_ops.py -m gpu -q -k "fused_operand or golden or euler or conv" > gpurun_out/ab6_parity.log 2>&1
echo "== parity exit $?"; tail -2 gpurun_out/ab6_parity.log | cut -c1-200
run() {  # label lib batch
  FLOWSE_LIB=$PWD/flowmse_b200/$2 timeout 600 python bench.py --steps 10 --batch $3 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/ab6.json 2> gpurun_out/ab6.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab6.json"))
print("$1 B=$3: value",round(d["value"]),"ms",round(d["ms_per_step"],3), "halo", d["roofline"]["nfe_ms_by_kernel_family"]["conv_halo"], "frac", round(d["roofline"]["frac"],4))
PY
}
for rep in 1 2 3; do
  run "old" libflowse_old.so 1
  run "new" libflowse.so 1
done
run "old" libflowse_old.so 8
run "new" libflowse.so 8
