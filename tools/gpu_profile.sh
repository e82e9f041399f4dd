#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time of one steady-state sampler step, (2) full capture of the
# dominant kernel.  Usage: tools/gpu_profile.sh <tag> [kernel-regex]
tag=${1:-r1}
kre=${2:-conv_gemm_tcgen05}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3400 --launch-count 1300 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list: $(wc -l < gpurun_out/launches_$tag.csv) lines"
ncu --set full --clock-control none --import-source on -k regex:$kre --launch-skip 105 --launch-count 3 \
    -o gpurun_out/prof_$tag -f python tools/run_nfe.py > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
