"""BASELINE configs[3]: one large pair, source points sharded over the ranks, target (and its octree) replicated; the
24 moments are exchanged peer-to-peer from inside the evaluation kernel every LM iteration.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/shard_bench.py [rings] [az] [reps]

Defaults are the 10M-point pair (320 rings x 31250 azimuths).  World size 1 runs the same pair on one GPU.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from probabilistic_point_clouds_registration_b200 import capi, multi, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rings = int(sys.argv[1]) if len(sys.argv) > 1 else 320
az = int(sys.argv[2]) if len(sys.argv) > 2 else 31250
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t0 = time.perf_counter()
src, tgt, _ = synth.lidar_pair(4, rings, az)
src, tgt = np.ascontiguousarray(src), np.ascontiguousarray(tgt)
gen_s = time.perf_counter() - t0
params = capi.make_params(max_neighbours=10, radius=0.5, dof=5.0)
d_tgt = torch.from_numpy(tgt).cuda()
stages = bool(os.environ.get("SHARD_STAGES"))
opt = capi.make_options(device=local, input_on_device=True, driver=1 if stages else 0, record_stage_times=stages)
# how the source is dealt to the ranks: contig (slice_bounds), strided (rank::world), block:<points> (block-cyclic)
# SHARD_FAKE="rank/world": pick that rank's share of the source but run it alone (world 1): what a share costs without any exchange
fake = os.environ.get("SHARD_FAKE")
d_rank, d_world = (int(v) for v in fake.split("/")) if fake else (rank, world)
for mode in os.environ.get("SHARD_MODES", "contig").split(","):
    if mode == "contig":
        lo, hi = multi.slice_bounds(len(src), d_rank, d_world)
        mine = src[lo:hi]
    elif mode == "strided":
        mine = np.ascontiguousarray(src[d_rank::d_world])
    elif mode.startswith("morton:"):
        mine = np.ascontiguousarray(src[multi.morton_chunk_indices(src, d_rank, d_world, int(mode.split(":")[1]))])
    else:
        mine = np.ascontiguousarray(src[multi.block_cyclic_indices(len(src), d_rank, d_world, int(mode.split(":")[1]))])
    d_src = torch.from_numpy(mine).cuda()
    times = []
    for rep in range(reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with multi.ShardedRegistration(d_src.data_ptr(), d_tgt.data_ptr(), params, rank, world, opt, n_source=len(mine),
                                       n_target=len(tgt)) as reg:
            reg.align()
            stats = reg.iteration_stats()
            hist = reg.transformation_history()
            ex = reg.stage_times()
            if rep >= 1:
                print(f"[rank {rank}] {mode} rep {rep}: {1e3 * (time.perf_counter() - t0):.1f} ms so far; {ex.exchanges} exchanges, {ex.exchange_wait_ms:.2f} ms in them "
                      f"({1e3 * ex.exchange_wait_ms / max(ex.exchanges, 1):.1f} us each, waiting for the slowest rank included)", flush=True)
            if stages:
                lt = reg.stage_times()
                print(f"[rank {rank}] {mode} rep {rep}: search {lt.search_ms:.1f} ms in {lt.search_launches} launches, eval {lt.eval_ms:.1f} ms in "
                      f"{lt.eval_launches} launches, K {sum(s['n_correspondences'] for s in stats)}", flush=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if rep > 0:
            times.append(dt)
    if rank == 0:
        corr = sum(s["n_correspondences"] for s in stats)
        evals = sum(s["lm_iterations"] + 1 for s in stats)
        ms = 1e3 * float(np.median(times))
        print(f"SHARD_BENCH {mode} n_gpus={world} n_src={len(src)} outer={len(stats)} evals={evals} correspondences={corr} "
              f"median ms={ms:.1f} reps: " + " ".join(f"{1e3 * t:.1f}" for t in times), flush=True)
if world > 1:
    dist.destroy_process_group()
