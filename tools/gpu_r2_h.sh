#!/bin/bash
# Round 2, call H: rolled row-stream transforms (halo + per-tap): parity, counters, A/B bench; RK45 diagnosis.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_parity_r2.py tests/test_gpu_configs.py -m gpu -q -s -k "not config3 and not three_way and not black_box" > gpurun_out/h_fwd.log 2>&1; echo "fwd exit $?"; tail -3 gpurun_out/h_fwd.log | cut -c1-300; grep "launches per evaluation" gpurun_out/h_fwd.log | tail -3
FLOWSE_FUSE_PREP=1 timeout 300 python tools/xf_diag.py 2>&1 | tail -23
FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/h_dbg.txt > /dev/null
grep "halo dbg" gpurun_out/h_dbg.txt | head -8 | cut -c1-420
grep "conv dbg" gpurun_out/h_dbg.txt | sed -n '1,4p;30,33p' | cut -c1-300
for fuse in 0 1 2; do
  FLOWSE_FUSE_PREP=$fuse timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/h_bench_f$fuse.json 2> gpurun_out/h_bench_f$fuse.err
  python - <<PY
import json
d=json.load(open("gpurun_out/h_bench_f$fuse.json"))
print("fuse $fuse: value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"],d["roofline"]["nfe_ms_by_kernel_family"])
PY
done
timeout 600 python tools/dbg_rk45.py 2>&1 | tail -12
