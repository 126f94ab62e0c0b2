TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/shard_bench.py 320 31250 6 > gpurun_out/shard2_contig.log 2>&1
PPCR_DRIVER=1 $TR tools/shard_bench.py 320 31250 6 > gpurun_out/shard2_contig_host.log 2>&1
SHARD_STRIDED=1 $TR tools/shard_bench.py 320 31250 4 > gpurun_out/shard2_strided.log 2>&1
SHARD_STAGES=1 $TR tools/shard_bench.py 320 31250 2 > gpurun_out/shard2_contig_stages.log 2>&1
grep -h "SHARD_BENCH\|search " gpurun_out/shard2_*.log
python tools/c4_parity.py 3 > gpurun_out/c4_parity.log 2>&1; cat gpurun_out/c4_parity.log
