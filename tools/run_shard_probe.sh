N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SHARD_MODES=contig,block:512,block:2048,block:8192,block:32768,strided $TR tools/shard_bench.py 320 31250 3 > gpurun_out/shard${N}_modes.log 2>&1
SHARD_STAGES=1 SHARD_MODES=block:2048 $TR tools/shard_bench.py 320 31250 1 > gpurun_out/shard${N}_block_stages.log 2>&1
grep -h "SHARD_BENCH\|search " gpurun_out/shard${N}_*.log
