#!/bin/bash
# What the driver runs at round end, in one call: the GPU test suite, smoke(), the default bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["clocks"])
print({k:r[k] for k in ("achieved","frac","issued_frac","traffic","ncu_tensor_pipe_active_pct_of_elapsed")}, r["kernel_alone"].get("frac"), d["cpu_baseline"]["value"])
PY
