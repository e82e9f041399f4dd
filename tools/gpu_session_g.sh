#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_evaluate.py -m gpu -x -q 2>&1 | tail -15
echo "== evaluate CLI, 48 synthetic utterances, N=5, 1 GPU"
timeout 900 python -m flowmse_b200.evaluate --folder_destination /tmp/eval1 --synthetic_utts 48 --synthetic_weights 0 --N 5 --seed 0 2>&1 | tail -3
cp /tmp/eval1/_timing.json gpurun_out/eval_synth48_n1.json
cat /tmp/eval1/_avg_results.txt
