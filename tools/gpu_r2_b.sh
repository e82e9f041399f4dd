#!/bin/bash
# Round 2, call B: fused operand prep (XF halo kernel) - parity first, then A/B bench.
mkdir -p gpurun_out
for f in tests/test_gpu_parity_r2.py tests/test_gpu_forward.py tests/test_gpu_configs.py tests/test_gpu_ops.py; do
  name=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q -s > gpurun_out/$name.log 2>&1
  echo "== $f: exit $?"; tail -3 gpurun_out/$name.log | cut -c1-300
done
grep -h "parity_r2\] fuse_prep" gpurun_out/test_gpu_parity_r2.log
for fuse in 0 1; do
  FLOWSE_FUSE_PREP=$fuse timeout 600 python bench.py --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/r2b_bench_fuse$fuse.json 2> gpurun_out/r2b_bench_fuse$fuse.err
  echo "fuse=$fuse exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2b_bench_fuse$fuse.json"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"],d["roofline"]["nfe_ms_by_kernel_family"], "halo frac", round(d["roofline"]["frac"],4))
PY
done
for b in 4 8; do
  timeout 600 python bench.py --no-cpu-baseline --no-torch-reference --config4 0 --batch $b > gpurun_out/r2b_bench_b$b.json 2> gpurun_out/r2b_bench_b$b.err
  python -c "
import json
d=json.load(open('gpurun_out/r2b_bench_b$b.json')); print('B=$b value',round(d['value']),'e2e',round(d['e2e']['value']))"
done
