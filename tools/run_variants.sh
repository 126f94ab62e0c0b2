# usage: tools/run_variants.sh name1 name2 ...   (csrc/tune_<name>.so; "default" = the product library)
for v in "$@"; do
  if [ "$v" = default ]; then unset PPCR_CUDA_LIB; else export PPCR_CUDA_LIB=probabilistic_point_clouds_registration_b200/csrc/tune_$v.so; fi
  echo "=== $v"
  C4_ITERS=12 python tools/c4_probe.py "" 2>&1 | grep "rep 1"
  python tools/run_once.py c3 1000 1 2>&1 | grep "rep 1" | sed 's/; launches.*//'
done
