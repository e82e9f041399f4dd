#!/bin/bash
# Round 2, call T: the whole GPU suite, the default bench line, the reference arm.
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.json
for f in tests/test_gpu_*.py; do
  name=$(basename $f .py)
  timeout 1500 python -m pytest $f -m gpu -q -s > gpurun_out/t_$name.log 2>&1
  echo "== $f: exit $?"; tail -1 gpurun_out/t_$name.log | cut -c1-200
done
timeout 900 python bench.py > gpurun_out/t_bench_n1.json 2> gpurun_out/t_bench_n1.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t_bench_reference.json 2> gpurun_out/t_bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/t_bench_n1.json"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"])
print("roofline frac",d["roofline"]["frac"],"alone",d["roofline"]["kernel_alone"]["frac"],"prep variant",d["roofline"]["standalone_prep_variant"]["frac"])
print("torch",d["gpu_torch_reference"])
print("cpu",d["cpu_baseline"]); print("extra",d["extra"]); print("clocks",d["clocks"])
r=json.load(open("gpurun_out/t_bench_reference.json")); print("ref",r["value"],r["config"],r["cpu_baseline"])
PY
