#!/bin/bash
# Round 2, call C: fused attention parity, XF diagnostics (role layouts, wait counters), quick A/B bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k attention > gpurun_out/c_attn.log 2>&1; echo "attn op exit $?"; tail -3 gpurun_out/c_attn.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_parity_r2.py -m gpu -q -k "not config3 and not three_way" > gpurun_out/c_fwd.log 2>&1; echo "fwd exit $?"; tail -3 gpurun_out/c_fwd.log | cut -c1-300
for lay in 0 1; do
  echo "=== layout $lay"
  FLOWSE_HALO_LAYOUT=$lay timeout 300 python tools/xf_diag.py 2>&1 | tail -22
done
FLOWSE_HALO_LAYOUT=0 FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/c_dbg_layout0.txt > /dev/null
FLOWSE_HALO_LAYOUT=1 FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/c_dbg_layout1.txt > /dev/null
grep "halo dbg" gpurun_out/c_dbg_layout0.txt | head -12
echo ...; grep "halo dbg" gpurun_out/c_dbg_layout1.txt | head -12
for cfg in "0 0" "0 1" "1 0" "1 1"; do
  set -- $cfg
  FLOWSE_HALO_LAYOUT=$1 FLOWSE_FUSE_PREP=$2 timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/c_bench_l$1_f$2.json 2> gpurun_out/c_bench_l$1_f$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/c_bench_l$1_f$2.json"))
print("layout $1 fuse $2: value",round(d["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"],d["roofline"]["nfe_ms_by_kernel_family"])
PY
done
