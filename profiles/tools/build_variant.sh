#!/bin/bash
mkdir -p variants
# usage (from the repo root): profiles/tools/build_variant.sh NAME [extra nvcc -D flags...]   -> variants/NAME.so
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart shared \
  -Xlinker -rpath,/usr/local/cuda/lib64 "$@" -o variants/$name.so tetrad_b200/csrc/tq_kernels.cu tetrad_b200/csrc/tq_api.cu tetrad_b200/csrc/tq_sink.cpp 2>&1 | grep -v "^$" | head -5
