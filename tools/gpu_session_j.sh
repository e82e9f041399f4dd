#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for mode in 2 3; do
FLOWSE_PDL=$mode timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_j_$mode.json 2> gpurun_out/bench_j_$mode.err
python - $mode <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_j_{m}.json").read().strip().splitlines()[-1])
    print("pdl", m, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("pdl", m, "FAILED", open(f"gpurun_out/bench_j_{m}.err").read()[-600:])
PY
done
done
FLOWSE_PDL=3 timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -3
