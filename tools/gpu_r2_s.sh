#!/bin/bash
# Round 2, call S: attention CTA-stagger experiment (FLOWSE_ATTN_ROT) + rows-per-CTA choice of the QKV kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k attention > gpurun_out/s_ops.log 2>&1; echo "ops exit $?"; tail -2 gpurun_out/s_ops.log
for rot in 0 1; do
  FLOWSE_ATTN_ROT=$rot timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k attention > gpurun_out/s_ops_rot$rot.log 2>&1; echo "rot $rot ops exit $?"
  FLOWSE_ATTN_ROT=$rot timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 196 --launch-count 196 -k regex:attn --csv --log-file gpurun_out/s_launches_rot$rot.csv python tools/run_nfe.py 3 0 > gpurun_out/s_ncu.log 2>&1
  python - <<PY
import csv,re,collections
lines=[l for l in open('gpurun_out/s_launches_rot$rot.csv') if l.startswith('"')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for r in rows:
    n=re.sub(r'\(.*','',r['Kernel Name']).replace('flowse::<unnamed>::','')
    k=(n[:40],r['Grid Size']); a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r['Metric Value'])/1e3
for k,(c,t) in agg.items(): print(f"rot $rot {t:8.1f} us {c:3d} x {t/c:7.1f} {k}")
PY
  FLOWSE_ATTN_ROT=$rot timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/s_bench_rot$rot.json 2> gpurun_out/s_bench_rot$rot.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s_bench_rot$rot.json"))
print("rot $rot: value",round(d["value"]),"ms",round(d["ms_per_step"],3), d["roofline"].get("nfe_ms_by_kernel_family"))
PY
done
