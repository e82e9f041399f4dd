#!/bin/bash
# Round 2, call R: attention rows-per-CTA variants + STFT transform/window variants; quick bench; launch list of one NFE.
mkdir -p gpurun_out
for f in tests/test_gpu_stft.py tests/test_gpu_ops.py tests/test_gpu_forward.py; do
  name=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q -s > gpurun_out/r_$name.log 2>&1
  echo "== $f: exit $?"; tail -3 gpurun_out/r_$name.log | cut -c1-300
done
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "euler5 or whole or raw_v or fused" > gpurun_out/r_parity.log 2>&1
echo "== parity subset: exit $?"; tail -3 gpurun_out/r_parity.log | cut -c1-300
for b in 1 4; do
  timeout 600 python bench.py --steps 10 --batch $b --no-cpu-baseline --no-torch-reference --config4 0 > gpurun_out/r_bench_b$b.json 2> gpurun_out/r_bench_b$b.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r_bench_b$b.json"))
print("B=$b: value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"], d["roofline"].get("nfe_ms_by_kernel_family"))
print(" standalone_prep_variant:", d["roofline"].get("standalone_prep_variant"))
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 196 --launch-count 196 --csv --log-file gpurun_out/r_launches.csv python tools/run_nfe.py 2 0 > gpurun_out/r_ncu.log 2>&1
python - <<'PY'
import csv,re,collections
lines=[l for l in open('gpurun_out/r_launches.csv') if l.startswith('"')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict(); tot=0
for r in rows:
    n=re.sub(r'\(.*','',r['Kernel Name']).replace('flowse::<unnamed>::','')
    if 'attn' not in n: continue
    k=(n[:40],r['Grid Size']); a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r['Metric Value'])/1e3
for k,(c,t) in agg.items(): print(f"{t:8.1f} us {c:3d} x {t/c:7.1f} {k}")
PY
