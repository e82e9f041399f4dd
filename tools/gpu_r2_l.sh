#!/bin/bash
# Round 2, call L: NB=3 vs NB=6 rows per straight-line batch.
mkdir -p gpurun_out
for lib in libflowse.so libflowse_nb6.so; do
  echo "=== $lib"
  export FLOWSE_LIB=$PWD/flowmse_b200/$lib
  FLOWSE_FUSE_PREP=1 timeout 300 python tools/xf_diag.py 2>&1 | sed -n '2,8p'
  FLOWSE_FUSE_PREP=1 FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/l_dbg.txt > /dev/null
  grep "halo dbg XF" gpurun_out/l_dbg.txt | sed -n '1,2p;6,7p' | cut -c200-420
  FLOWSE_FUSE_PREP=1 timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/l_bench.json"))
print("$lib fuse 1: value",round(d["value"]),"ms",round(d["ms_per_step"],3),d["roofline"]["nfe_ms_by_kernel_family"])
PY
done
