#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -5
echo "== dbg conv"
timeout 300 python tools/dbg_conv.py 2>&1 | grep "conv dbg\|---" | tail -40
echo "== bench"
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_b.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["roofline"]["nfe_ms_by_kernel_family"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 258 --launch-count 262 --csv \
    --log-file gpurun_out/launches_b.csv python tools/run_nfe.py 2 0 > gpurun_out/ncu_launch.log 2>&1
grep -c head_conv gpurun_out/launches_b.csv; grep head_conv gpurun_out/launches_b.csv | awk -F'","' '{print $5, $9, $NF}' | cut -c1-150
