#!/bin/bash
# registers / spills / stack per kernel of libppcr_cuda.so (cross-compiles, no GPU needed).  Extra nvcc flags may follow.
cd "$(dirname "$0")/../probabilistic_point_clouds_registration_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I ../../include \
     -Xptxas=-v "$@" -c ppcr_capi.cu -o /tmp/ppcr_ptxas_report.o 2>&1 |
  sed 's/ptxas info    : //' |
  awk '/Compiling entry function/{name=$4} /bytes stack frame/{sp=$0} /Used [0-9]+ registers/{print name " | " sp " | " $0}' |
  grep -v cub | c++filt | cut -c1-320
