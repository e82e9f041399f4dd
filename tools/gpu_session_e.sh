#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for mode in 0 2 1; do
FLOWSE_PDL=$mode timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_e_$mode.json 2> gpurun_out/bench_e_$mode.err
python - $mode <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_e_{m}.json").read().strip().splitlines()[-1])
    print("pdl", m, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("pdl", m, "FAILED", open(f"gpurun_out/bench_e_{m}.err").read()[-500:])
PY
done
done
FLOWSE_PDL=2 timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -3
for b in 2 4 8; do
timeout 600 python bench.py --steps 5 --batch $b --no-cpu-baseline > gpurun_out/bench_e_b$b.json 2> gpurun_out/bench_e_b$b.err
python - $b <<'PY'
import json,sys
m=sys.argv[1]
d=json.loads(open(f"gpurun_out/bench_e_b{m}.json").read().strip().splitlines()[-1])
print("batch", m, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["nfe_ms_by_kernel_family"])
PY
done
