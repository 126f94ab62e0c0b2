python -m pytest tests/test_gpu_search.py tests/test_gpu_align.py -m gpu -x -q 2>&1 | tail -4
for f in 0 4; do echo "== PPCR_Q_FLAGS=$f"; PPCR_Q_FLAGS=$f python tools/run_once.py c3 1000 1 2>&1 | grep "rep 1" | sed 's/; launches.*//'; PPCR_Q_FLAGS=$f C4_ITERS=12 python tools/c4_probe.py "" 2>&1 | grep "rep 1"; done
PPCR_Q_FLAGS=0 python tools/batch_bench.py 96 6 | tail -1
PPCR_Q_FLAGS=4 python tools/batch_bench.py 96 6 | tail -1
