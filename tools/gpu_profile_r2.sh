#!/bin/bash
# Round-2 evidence for profiles/: bench lines, launch list of one NFE, ncu --set full of every kernel family.
# The .ncu-rep files are summarised ON THE BOX (tools/ncu_summary.py) and deleted: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
summarise() {  # name
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/$1_summary.json && rm -f gpurun_out/$1.ncu-rep
  python -c "import json;d=json.load(open('gpurun_out/$1_summary.json'));print('$1:',len(d),'launches summarised')"
}
echo "== launch list of one NFE (eager, second evaluation; B=1, T=512)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 196 --launch-count 196 --csv \
    --log-file gpurun_out/r2_launches_one_nfe.csv python tools/run_nfe.py 2 0 > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/r2_launches_one_nfe.csv
echo "== ncu full: halo kernels of the second NFE"
timeout 1200 ncu --set full --clock-control none -k regex:conv_halo_kernel --launch-skip 38 --launch-count 38 \
    -o gpurun_out/r2_conv_halo_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_halo.log 2>&1
summarise r2_conv_halo_full
echo "== ncu full: the same layers with the standalone prep pass (fuse_prep = 0): halo kernel + prep"
FLOWSE_FUSE_PREP=0 timeout 1200 ncu --set full --clock-control none -k regex:"conv_halo_kernel|gn_prep_plain" --launch-skip 126 --launch-count 40 \
    -o gpurun_out/r2_unfused_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_unfused.log 2>&1
summarise r2_unfused_full
echo "== ncu full: attention, pyramid heads, input conv, combine, FIR, final, resampling prep, time embedding"
timeout 1200 ncu --set full --clock-control none \
    -k regex:"attn_|head_conv|conv_in|combine|fir_down4|final_kernel|gn_prep_resample|temb_" --launch-skip 43 --launch-count 43 \
    -o gpurun_out/r2_small_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_small.log 2>&1
summarise r2_small_full
echo "== ncu full: low-resolution per-tap conv + its prep (a sample)"
timeout 900 ncu --set full --clock-control none -k regex:"conv_gemm_tcgen05|gn_prep_plain" --launch-skip 130 --launch-count 16 \
    -o gpurun_out/r2_lowres_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_lowres.log 2>&1
summarise r2_lowres_full
echo "== final kernel (fused Euler update) + prior at B=16: 2 Mi bins"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
    -k regex:"final_kernel|prior_kernel" --csv --log-file gpurun_out/r2_final_b16.csv python tools/run_sample.py 16 512 2 > gpurun_out/ncu_final.log 2>&1
echo "== ncu full: STFT / iSTFT kernels"
timeout 600 ncu --set full --clock-control none --launch-skip 20 --launch-count 10 \
    -o gpurun_out/r2_stft_full -f python tools/run_stft.py > gpurun_out/ncu_full_stft.log 2>&1
summarise r2_stft_full
echo "== source-level capture of ONE fused halo launch (kept as .ncu-rep)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel --launch-skip 41 --launch-count 1 \
    -o gpurun_out/r2_halo_xf_source -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_src.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== bench n=1 (default flags)"
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.json; echo
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cut -c1-300 gpurun_out/r2_bench_reference.json
du -sh gpurun_out
