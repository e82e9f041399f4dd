#!/bin/bash
# Round 2, call J: rolled transform with 3 vs 6 rows in flight; RK45 + packed tests.
mkdir -p gpurun_out
for ah in 3 6; do
  echo "=== FLOWSE_XF_AHEAD=$ah"
  FLOWSE_XF_AHEAD=$ah FLOWSE_FUSE_PREP=1 timeout 300 python tools/xf_diag.py 2>&1 | sed -n '1,8p'
  FLOWSE_XF_AHEAD=$ah FLOWSE_FUSE_PREP=1 FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/j_dbg_$ah.txt > /dev/null
  grep "halo dbg XF" gpurun_out/j_dbg_$ah.txt | sed -n '1,2p;6,7p' | cut -c1-420
  FLOWSE_XF_AHEAD=$ah FLOWSE_FUSE_PREP=1 timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/j_bench_a$ah.json 2> gpurun_out/j_bench_a$ah.err
  python - <<PY
import json
d=json.load(open("gpurun_out/j_bench_a$ah.json"))
print("ahead $ah fuse 1: value",round(d["value"]),"ms",round(d["ms_per_step"],3),d["roofline"]["nfe_ms_by_kernel_family"])
PY
done
timeout 1200 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "black_box" > gpurun_out/j_n4.log 2>&1; echo "n4 exit $?"; tail -3 gpurun_out/j_n4.log | cut -c1-300; grep "parity_r2\]" gpurun_out/j_n4.log | cut -c1-400
