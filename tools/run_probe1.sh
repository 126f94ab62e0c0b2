bash tools/ncu_capture.sh r02 "4 1 0" c3 > /dev/null 2>&1
python tools/time_kernels.py c3 > gpurun_out/r02_time_kernels.txt 2>&1
PPCR_DRIVER=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu --headline-only > gpurun_out/r02_launches_bench.log 2>&1
cat gpurun_out/r02_time_kernels.txt
