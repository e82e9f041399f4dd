#!/bin/bash
# compute-sanitizer memcheck over one small NCSN++ evaluation (T=64, B=2) and the STFT round trip
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/run_nfe.py 1 0 2 64 > gpurun_out/sanitize_nfe.log 2>&1; echo "memcheck nfe rc=$?"; tail -5 gpurun_out/sanitize_nfe.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_stft.py -m gpu -x -q -k "golden or ragged" > gpurun_out/sanitize_stft.log 2>&1; echo "memcheck stft rc=$?"; tail -5 gpurun_out/sanitize_stft.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/run_nfe.py 1 0 1 64 > gpurun_out/racecheck_nfe.log 2>&1; echo "racecheck nfe rc=$?"; tail -8 gpurun_out/racecheck_nfe.log
