N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SHARD_MODES=block:8192 $TR tools/shard_bench.py 320 31250 6 > gpurun_out/shard${N}_reps.log 2>&1
SHARD_STAGES=1 SHARD_MODES=block:8192 $TR tools/shard_bench.py 320 31250 3 > gpurun_out/shard${N}_stages.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv >> gpurun_out/shard${N}_reps.log
grep -h "SHARD_BENCH\|rank" gpurun_out/shard${N}_reps.log gpurun_out/shard${N}_stages.log | sort -k4,4 -k2,2 | head -80
