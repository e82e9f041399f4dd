#!/bin/bash
# Round 2, call U (N GPUs): bench.py under torchrun - weak-scaling headline + configs[3] strong-scaling leg.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/u_bench_n$N.json 2> gpurun_out/u_bench_n$N.err
echo "bench exit $?"
tail -3 gpurun_out/u_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/u_bench_n$N.json") if l.startswith("{")][-1])   # NCCL may print its version line first
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"per-rank",d.get("per_rank_ms_per_step"))
print("extra",d["extra"])
PY
[ "$2" = "noref" ] && exit 0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/u_ref_n$N.json 2> gpurun_out/u_ref_n$N.err
echo "ref exit $?"; cat gpurun_out/u_ref_n$N.json | cut -c1-300
