#!/bin/bash
# ncu --set full of the HBM-bound kernels: the Euler update (768 MiB working set) and the top-level operand prep
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:axpy_kernel --launch-skip 2 --launch-count 2 \
    -o gpurun_out/r1b_euler_full -f python tools/run_euler.py > gpurun_out/ncu_full_euler.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_prep_plain --launch-skip 88 --launch-count 3 \
    -o gpurun_out/r1b_prep_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_prep.log 2>&1
for f in r1b_euler_full r1b_prep_full; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.csv 2>/dev/null
done
ls -la gpurun_out/r1b_euler_full.* gpurun_out/r1b_prep_full.*
