#!/bin/bash
# One `ncu --set full` capture per hot kernel of a workload (run under gpurun; results in gpurun_out/).
#   tools/ncu_capture.sh [tag] [kernels: "4 1 0"] [workload]
tag=${1:-cap}
kernels=${2:-"4 1 0"}
wl=${3:-c3}
mkdir -p gpurun_out
for which in $kernels; do
  PPCR_PROFILE_KERNEL=$which timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -f -o gpurun_out/${tag}_k${which} python tools/time_kernels.py $wl 1000 1 > gpurun_out/${tag}_k${which}.log 2>&1
done
ls -la gpurun_out/ | grep ${tag}
