#!/bin/bash
# Development aid: build a variant of libwofdm.so that differs in the macros of ONE translation unit.
#   tools/build_variant.sh <tag> <source.cu> [-DNAME=VALUE ...]   ->  build/variants/libwofdm_<tag>.so   (run with WOFDM_LIB=...)
set -e
cd "$(dirname "$0")/../w-ofdm-optimization_b200/csrc"
tag=$1; src=$2; shift 2
mkdir -p ../../build/variants ../../build/obj
obj=../../build/variants/${src%.cu}_$tag.o
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr "$@" -c $src -o $obj
others=$(ls ../../build/obj/*.o | grep -v "/${src%.cu}.o" | grep -v "_dbg.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/libwofdm_$tag.so $others $obj -lcudart
echo built build/variants/libwofdm_$tag.so
