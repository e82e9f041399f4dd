#!/bin/bash
mkdir -p gpurun_out
FLOWSE_SPLITK=l2 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -4
echo "== dbg conv (l2)"
FLOWSE_SPLITK=l2 timeout 300 python tools/dbg_conv.py 2>&1 | grep "conv dbg\|---" | tail -19 | awk 'NR%3==2' | cut -c1-220
for rep in 1 2; do
for mode in dsmem l2; do
FLOWSE_SPLITK=$mode timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_i_$mode.json 2> gpurun_out/bench_i_$mode.err
python - $mode <<'PY'
import json,sys
m=sys.argv[1]
d=json.loads(open(f"gpurun_out/bench_i_{m}.json").read().strip().splitlines()[-1])
print(m, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"])
PY
done
done
