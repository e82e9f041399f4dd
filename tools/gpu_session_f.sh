#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stft.py -m gpu -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -5
