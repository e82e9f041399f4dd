#!/bin/bash
# Round-2 evidence refresh after the last kernel changes (512-thread fused halo kernel, attention rows per CTA):
# launch list of one NFE + ncu --set full of the halo launches and of the small kernel families.
mkdir -p gpurun_out
summarise() {  # name
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/$1_summary.json && rm -f gpurun_out/$1.ncu-rep
  python -c "import json;d=json.load(open('gpurun_out/$1_summary.json'));print('$1:',len(d),'launches summarised')"
}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 196 --launch-count 196 --csv \
    --log-file gpurun_out/r2_launches_one_nfe.csv python tools/run_nfe.py 2 0 > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/r2_launches_one_nfe.csv
timeout 1200 ncu --set full --clock-control none -k regex:conv_halo_kernel --launch-skip 38 --launch-count 38 \
    -o gpurun_out/r2_conv_halo_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_halo.log 2>&1
summarise r2_conv_halo_full
timeout 1200 ncu --set full --clock-control none \
    -k regex:"attn_|head_conv|conv_in|combine|fir_down4|final_kernel|gn_prep_resample|temb_" --launch-skip 43 --launch-count 43 \
    -o gpurun_out/r2_small_full -f python tools/run_nfe.py 2 0 > gpurun_out/ncu_full_small.log 2>&1
summarise r2_small_full
du -sh gpurun_out
