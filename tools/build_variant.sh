#!/bin/sh
# build a tuning variant of libmle_cuda.so: tools/build_variant.sh <name> <extra nvcc flags...>  ->  tools/bin/libmle_<name>.so
# (select it at run time with MLE_CUDA_LIB=tools/bin/libmle_<name>.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/bin
nvcc -gencode arch=compute_100a,code=sm_100a "$@" -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
     -o tools/bin/libmle_$name.so multilevelestimators.jl_b200/csrc/mle_cuda.cu
echo built tools/bin/libmle_$name.so
