#!/bin/bash
# Build A/B variants of libnphusl_cuda.so with different launch shapes (measurement tool).
# usage: build_variants.sh name1 "-DFLAGS ..." name2 "-DFLAGS ..." ...   -> build/var/lib_<name>.so
set -e
cd "$(dirname "$0")/../../husl-numpy_b200/csrc"
mkdir -p ../../build/var
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xptxas -v \
      -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2 --cudart static $flags -shared nphusl_cuda_api.cu \
      -o ../../build/var/lib_$name.so 2> ../../build/var/$name.log && echo "built $name" ) &
done
wait
