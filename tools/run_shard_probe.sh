#!/bin/bash
# The sharded 10M-point pair under different deals of the source (run under gpurun --gpus N):
#   tools/run_shard_probe.sh N "contig,block:8192,morton:4096"      modes: contig | strided | block:<points> | morton:<points>
# SHARD_STAGES=1 adds per-rank search / evaluation times (host-stepped driver); SHARD_FAKE=r/w runs rank r's share of w alone.
N=${1:-2}
MODES=${2:-contig,block:8192}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SHARD_MODES=$MODES $TR tools/shard_bench.py 320 31250 3 > gpurun_out/shard${N}_modes.log 2>&1
grep -h "SHARD_BENCH\|search " gpurun_out/shard${N}_modes.log
