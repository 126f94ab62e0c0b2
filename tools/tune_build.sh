#!/bin/bash
# Builds differently tuned copies of libppcr_cuda.so into gpurun_out-independent scratch (csrc/tune_*.so, git-ignored):
#   tools/tune_build.sh name "-DPPCR_EVAL_FAST_BLOCKS=5 -DPPCR_EVAL_BATCH=5"
# Use with PPCR_CUDA_LIB=probabilistic_point_clouds_registration_b200/csrc/tune_<name>.so
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -shared \
  -ccbin /usr/bin/g++ -I include $@ -Xptxas=-v -o probabilistic_point_clouds_registration_b200/csrc/tune_${name}.so \
  probabilistic_point_clouds_registration_b200/csrc/ppcr_capi.cu -lcudart 2>&1 | grep -A3 "k_evalctlILb1\|k_searchILi[04]" | grep "Used\|spill"
