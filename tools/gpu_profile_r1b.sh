#!/bin/bash
# Round-1 (second half) evidence for profiles/: bench lines, launch list of one NFE, ncu --set full of the dominant kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
echo "== bench n=1 (default flags)"
timeout 900 python bench.py > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err; tail -c 400 gpurun_out/r1b_bench_n1.json; echo
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b_bench_reference.json 2> gpurun_out/r1b_bench_reference.err; cat gpurun_out/r1b_bench_reference.json | cut -c1-400
echo "== launch list of one NFE (eager, second evaluation)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 240 --launch-count 240 --csv \
    --log-file gpurun_out/r1b_launches_one_nfe.csv python tools/run_nfe.py 2 0 > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/r1b_launches_one_nfe.csv
echo "== ncu full: halo kernel, last 12 launches of the second NFE (top-level up path)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel --launch-skip 58 --launch-count 12 \
    -o gpurun_out/r1b_conv_halo_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_halo.log 2>&1
echo "== ncu full: pyramid head + cluster split-K conv"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:head_conv_kernel --launch-skip 7 --launch-count 7 \
    -o gpurun_out/r1b_head_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_head.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 --launch-skip 60 --launch-count 6 \
    -o gpurun_out/r1b_conv_gemm_cluster_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
