python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
PPCR_TRACE=1 python tools/batch_bench.py 12 1 > gpurun_out/c5_trace_1lane.log 2>&1; tail -40 gpurun_out/c5_trace_1lane.log
python tools/c4_probe.py "" > gpurun_out/c4_probe.log 2>&1; tail -5 gpurun_out/c4_probe.log
