for d in 1 2 4 6; do echo "== PPCR_SEARCH_Q_BATCH_DIV=$d"; PPCR_SEARCH_Q_BATCH_DIV=$d python tools/batch_bench.py 256 6 6 8 2>&1 | tail -3; done
