#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -5
echo "== dbg conv"
timeout 300 python tools/dbg_conv.py 2>&1 | grep "conv dbg\|---" | tail -19 | cut -c1-220
echo "== bench"
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["roofline"]["nfe_ms_by_kernel_family"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 258 --launch-count 258 --csv \
    --log-file gpurun_out/launches_c.csv python tools/run_nfe.py 2 0 > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv,re,collections
rows=list(csv.reader(open('gpurun_out/launches_c.csv')))
hdr=None;data=[]
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
agg=collections.OrderedDict()
for d in data:
    n=d['Kernel Name']; n=re.sub(r'flowse::<unnamed>::','',n); n=re.sub(r'^void ','',n).split('(')[0][:40]
    t=int(d['Metric Value'])/1000
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(a[1] for a in agg.values())
for n,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{n:42s} {a[0]:4d} {a[1]:8.1f} {100*a[1]/tot:5.1f}%")
print("total",round(tot,1), "launches", len(data))
PY
