python -m pytest tests/test_gpu_search.py tests/test_gpu_metrics.py -m gpu -x -q 2>&1 | tail -3
python tools/run_once.py c3 1000 1 2>&1 | grep "rep 1" | sed 's/; launches.*//'
python tools/run_once.py c3 1000 0 2>&1 | grep "rep 1" | sed 's/; 34 outer.*//'
C4_ITERS=12 python tools/c4_probe.py "" 2>&1 | grep "rep 1"
python tools/batch_bench.py 192 6 6 | tail -2
python tools/time_kernels.py c3 | head -4
