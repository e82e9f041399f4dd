#!/bin/bash
# Round 2, call W: programmatic dependent launch on the halo conv launches too (FLOWSE_PDL=4) vs the default (2).
mkdir -p gpurun_out
FLOWSE_PDL=4 timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_parity_r2.py -m gpu -q -k "euler or whole or golden" > gpurun_out/w_parity.log 2>&1
echo "== parity under PDL=4 exit $?"; tail -2 gpurun_out/w_parity.log | cut -c1-200
for rep in 1 2; do
for m in 2 4; do
  FLOWSE_PDL=$m timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/w_bench_pdl${m}_$rep.json 2> gpurun_out/w_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/w_bench_pdl${m}_$rep.json"))
print("pdl $m rep $rep: value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3))
PY
done
done
for m in 2 4; do
  FLOWSE_PDL=$m timeout 600 python bench.py --steps 10 --batch 4 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/w_bench_b4_pdl$m.json 2> gpurun_out/w_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/w_bench_b4_pdl$m.json"))
print("B=4 pdl $m: value",round(d["value"]),"ms",round(d["ms_per_step"],3))
PY
done
