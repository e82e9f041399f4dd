#!/bin/bash
# Round 2, call K: 3-row straight-line batches in the halo transform.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "fused" > gpurun_out/k_fwd.log 2>&1; echo "fused test exit $?"; tail -2 gpurun_out/k_fwd.log | cut -c1-300
FLOWSE_FUSE_PREP=1 timeout 300 python tools/xf_diag.py 2>&1 | tail -23
FLOWSE_FUSE_PREP=1 FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/k_dbg.txt > /dev/null
grep "halo dbg XF" gpurun_out/k_dbg.txt | sed -n '1,2p;6,7p' | cut -c1-420
for fuse in 0 1; do
  FLOWSE_FUSE_PREP=$fuse timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/k_bench_f$fuse.json 2> gpurun_out/k_bench_f$fuse.err
  python - <<PY
import json
d=json.load(open("gpurun_out/k_bench_f$fuse.json"))
print("fuse $fuse: value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"],d["roofline"]["nfe_ms_by_kernel_family"])
PY
done
