#!/bin/bash
# BASELINE.json configs[4]: VoiceBank-DEMAND-shaped synthetic test set (824 utterances), N in {1,5,25}, through the
# evaluate drop-in.  Usage: tools/config5_sweep.sh <n_gpus> <tag>
n=${1:-1}; tag=${2:-n1}
mkdir -p gpurun_out
for N in 1 5 25; do
  out=/tmp/eval_c5_${tag}_N$N
  if [ "$n" = "1" ]; then
    timeout 1500 python -m flowmse_b200.evaluate --folder_destination $out --synthetic_utts 824 --synthetic_weights 0 --N $N --seed 0 --max_batch_frames 4096 2>&1 | tail -1
  else
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + N)) -m flowmse_b200.evaluate --folder_destination $out --synthetic_utts 824 --synthetic_weights 0 --N $N --seed 0 --max_batch_frames 4096 2>&1 | grep frames_per_s | tail -1
  fi
  cp $out/_timing.json gpurun_out/${PREFIX:-r1b}_config5_${tag}_N$N.json
done
