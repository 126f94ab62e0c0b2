N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SHARD_MODES=$2 $TR tools/shard_bench.py 320 31250 3 > gpurun_out/shard${N}_morton.log 2>&1
grep -h "SHARD_BENCH" gpurun_out/shard${N}_morton.log
