#!/bin/bash
# Session A: correctness of the new kernels (cluster split-K, fused pyramid head) + A/B bench + launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_ops.py tests/test_gpu_forward.py; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -x -q 2>&1 | tail -30 > gpurun_out/$name.log
  echo "== $f: exit ${PIPESTATUS[0]}"; tail -8 gpurun_out/$name.log
done
echo "== bench default (cluster split-K)"
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_cluster.json 2> gpurun_out/bench_cluster.err; tail -c 600 gpurun_out/bench_cluster.json | head -c 300; echo
echo "== bench cluster max 8"
FLOWSE_CLUSTER16=0 timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_cluster8.json 2> gpurun_out/bench_cluster8.err
echo "== bench 2pass"
FLOWSE_SPLITK=2pass timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_2pass.json 2> gpurun_out/bench_2pass.err
python - <<'PY'
import json
for n in ("cluster","cluster8","2pass"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["roofline"]["nfe_ms_by_kernel_family"])
    except Exception as e:
        print(n, "FAILED", e, open(f"gpurun_out/bench_{n}.err").read()[-800:])
PY
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 --launch-count 320 --csv \
    --log-file gpurun_out/launches_r1b.csv python tools/run_nfe.py 2 0 > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/launches_r1b.csv
for f in tests/test_gpu_configs.py tests/test_gpu_torch_cuda.py; do
  name=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/$name.log
  echo "== $f: exit ${PIPESTATUS[0]}"; tail -8 gpurun_out/$name.log
done
