#!/bin/bash
# Round 2, call F: what bounds the halo transform (dbg modes), per-tap XF parity + bench, packed weights test.
mkdir -p gpurun_out
for mode in 0 1 2 4 3 7; do
  echo "--- FLOWSE_XF_DBGMODE=$mode (1 no loads, 2 no math, 4 no stores)"
  FLOWSE_XF_DBGMODE=$mode FLOWSE_FUSE_PREP=1 FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/f_dbg_$mode.txt > /dev/null
  grep "halo dbg XF" gpurun_out/f_dbg_$mode.txt | sed -n '1,2p;6,7p' | cut -c1-420
done
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_parity_r2.py tests/test_gpu_configs.py -m gpu -q -s -k "not config3 and not three_way" > gpurun_out/f_fwd.log 2>&1; echo "fwd exit $?"; tail -3 gpurun_out/f_fwd.log | cut -c1-300; grep "launches per evaluation" gpurun_out/f_fwd.log
FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/f_dbg.txt > /dev/null
grep "conv dbg" gpurun_out/f_dbg.txt | sed -n '1,4p;30,34p' | cut -c1-300
for fuse in 0 1 2; do
  FLOWSE_FUSE_PREP=$fuse timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/f_bench_f$fuse.json 2> gpurun_out/f_bench_f$fuse.err
  python - <<PY
import json
d=json.load(open("gpurun_out/f_bench_f$fuse.json"))
print("fuse $fuse: value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"],d["roofline"]["nfe_ms_by_kernel_family"])
PY
done
