run() { export PPCR_CUDA_LIB=probabilistic_point_clouds_registration_b200/csrc/tune_$1.so; shift; for b in "$@"; do echo -n "$PPCR_CUDA_LIB batch per_sm $b: "; PPCR_EVAL_PER_SM_BATCH=$b python tools/batch_bench.py 192 6 6 | tail -2 | tr '\n' ' '; echo; done; }
run old4 4 3
run la4 4 3 2
run la5 5 4 3
run la6 6 4
