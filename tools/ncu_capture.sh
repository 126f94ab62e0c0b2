#!/bin/bash
# One `ncu --set full` capture per hot kernel of the c3 workload (run under gpurun; results in gpurun_out/).
#   tools/ncu_capture.sh [tag]
tag=${1:-cap}
mkdir -p gpurun_out
for which in 4 1 0; do
  PPCR_PROFILE_KERNEL=$which timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -f -o gpurun_out/${tag}_k${which} python tools/time_kernels.py c3 1000 1 > gpurun_out/${tag}_k${which}.log 2>&1
done
ls -la gpurun_out/
