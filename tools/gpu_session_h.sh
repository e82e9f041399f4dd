#!/bin/bash
mkdir -p gpurun_out
echo "== bench 2 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.json | head -c 700; echo
echo "== evaluate 2 GPUs, 96 synthetic utterances"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m flowmse_b200.evaluate --folder_destination /tmp/eval2 --synthetic_utts 96 --synthetic_weights 0 --N 5 --seed 0 2>&1 | tail -4
cp /tmp/eval2/_timing.json gpurun_out/eval_synth96_n2.json
ls /tmp/eval2/files | wc -l
