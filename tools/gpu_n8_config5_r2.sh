#!/bin/bash
# BASELINE configs[4] (824 synthetic utterances) through the evaluate drop-in on 8 GPUs, N = 5 Euler, round 2.
mkdir -p gpurun_out
N=${1:-5}
out=/tmp/eval_c5_n8_N$N
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29540 + N)) -m flowmse_b200.evaluate --folder_destination $out --synthetic_utts 824 --synthetic_weights 0 --N $N --seed 0 --max_batch_frames 4096 2>&1 | grep frames_per_s | tail -1
cp $out/_timing.json gpurun_out/r2_config5_n8_N$N.json
