#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r1b_bench_n8.json 2> gpurun_out/r1b_bench_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/r1b_bench_n8.json').read().strip().splitlines()[-1])
print('bench n8', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
bash tools/config5_sweep.sh 8 n8
