#!/bin/bash
# A/B build of the product library with extra nvcc flags: tools/ab_build.sh NAME -DFOO=1 ...  ->  lib/libpamg_cuda_NAME.so
# (select it at run time with PAMG_LIB=p-a_multigrids_b200/lib/libpamg_cuda_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
P=p-a_multigrids_b200
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-O3,-Wall -shared -cudart static -I include -I $P/csrc "$@" \
  $P/csrc/pamg_api.cu $P/csrc/pamg_mesh.cpp $P/csrc/pamg_plan.cpp -o $P/lib/libpamg_cuda_$name.so -ldl
echo $P/lib/libpamg_cuda_$name.so
