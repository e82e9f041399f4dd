#!/bin/bash
# Build libflowse.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr"
mkdir -p build
pids=()
for f in engine conv_gemm conv_halo kernels_gn kernels_pointwise sgemm stft attention; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ flowse_internal.h -nt build/$f.o ] || [ ptx.cuh -nt build/$f.o ] || [ operand.cuh -nt build/$f.o ] || [ prep.cuh -nt build/$f.o ] || [ ../../include/flowse.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${XFLAGS} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -Wno-deprecated-gpu-targets -shared -o ../libflowse.so build/engine.o build/conv_gemm.o build/conv_halo.o build/kernels_gn.o build/kernels_pointwise.o build/sgemm.o build/stft.o build/attention.o -lcudart_static -lpthread -ldl -lrt
echo "built $(cd .. && pwd)/libflowse.so"
