#!/bin/bash
# Round 2, call G: transform timers; N3 / N4 tests.
mkdir -p gpurun_out
for mode in 0 7; do
  echo "--- FLOWSE_XF_DBGMODE=$mode"
  FLOWSE_XF_DBGMODE=$mode FLOWSE_FUSE_PREP=1 FLOWSE_CONV_DBG=1 timeout 300 python tools/run_nfe.py 1 0 2> gpurun_out/g_dbg_$mode.txt > /dev/null
  grep "halo dbg XF" gpurun_out/g_dbg_$mode.txt | sed -n '1,2p;6,7p' | cut -c250-520
done
timeout 1200 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "packed or rk_lincomb or black_box" > gpurun_out/g_n34.log 2>&1; echo "n3/n4 exit $?"; tail -5 gpurun_out/g_n34.log | cut -c1-400; grep "parity_r2\]" gpurun_out/g_n34.log | cut -c1-300
