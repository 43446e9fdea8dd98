#!/bin/bash
# build experiment variants of libgevb.so (extra -D flags) next to the product library: scripts/ablate.sh NAME "-DFLAG ..."
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/../gevolution-1.2_b200"
mkdir -p /tmp/abl_$NAME
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC --expt-relaxed-constexpr $FLAGS"
for f in ctx timing nccl_dl fft fourier_kernels source_kernels particles deposit geodesic spectrum; do /usr/local/cuda/bin/nvcc $NVF -c csrc/$f.cu -o /tmp/abl_$NAME/$f.o & done
/usr/local/cuda/bin/nvcc $NVF -x cu -c host/sim.cpp -o /tmp/abl_$NAME/sim.o &
wait
mkdir -p ../build/abl
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../build/abl/libgevb_$NAME.so /tmp/abl_$NAME/*.o -lcufft -ldl -Xlinker -rpath=/usr/local/cuda/lib64
echo built build/abl/libgevb_$NAME.so
