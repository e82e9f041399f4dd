#!/bin/bash
# Round 2, call I: ncu source-level profile of the XF halo kernel (what stalls the transform warps).
mkdir -p gpurun_out
FLOWSE_FUSE_PREP=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel -s 3 -c 2 -f -o gpurun_out/xf_prof python tools/run_nfe.py 1 0 > gpurun_out/i_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/i_ncu.log; ls -la gpurun_out/xf_prof.ncu-rep
