#!/bin/bash
mkdir -p gpurun_out
echo "--- debug lib, plain T=64 B=1"; FLOWSE_LIB=$PWD/flowmse_b200/libflowse_dbg.so timeout 300 python tools/run_nfe.py 1 0 1 64 > gpurun_out/o_dbg.log 2>&1; grep "mbar timeout" gpurun_out/o_dbg.log | sed 's/block [0-9]* //; s/bar 0x[0-9a-f]* //' | sort | uniq -c | sort -rn | head -30; tail -2 gpurun_out/o_dbg.log | cut -c1-200
