"""Worker of tests/test_gpu_multi.py (launched under torchrun, one rank per GPU): one pair sharded over the ranks,
every rank must end with the pose a single GPU computes."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import pose_delta  # noqa: E402
from probabilistic_point_clouds_registration_b200 import capi, multi, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {}
    for name, (src, tgt), kw, exact in (
            ("c1", synth.config1_plane_sphere()[:2], dict(max_neighbours=20, dof=5.0, radius=1.0), True),
            ("lidar", synth.lidar_pair(5, 48, 1500)[:2], dict(max_neighbours=10, dof=5.0, radius=0.5), True),
            # the default (float32-row, bulk-copy staged) evaluation: rows are identical on every rank, only the float64
            # summation order across rows differs from the single-GPU run
            ("lidar_default", synth.lidar_pair(6, 48, 1500)[:2], dict(max_neighbours=10, dof=5.0, radius=0.5), False)):
        params = capi.make_params(**kw)
        lo, hi = multi.slice_bounds(len(src), rank, world)
        opt = capi.make_options(device=local, exact_weights=exact)
        with multi.ShardedRegistration(src[lo:hi], tgt, params, rank, world, opt) as reg:
            reg.align()
            hist = reg.transformation_history()
            stats = reg.iteration_stats()
        # the single-GPU answer, computed on every rank
        with capi.Registration(src, tgt, params, capi.make_options(device=local, exact_weights=exact)) as one:
            one.align()
            ref = one.transformation_history()
            ref_stats = one.iteration_stats()
        assert len(hist) == len(ref), (name, len(hist), len(ref))
        assert [s["lm_iterations"] for s in stats] == [s["lm_iterations"] for s in ref_stats], name
        rot, tr = pose_delta(hist[-1], ref[-1])
        if exact:
            assert [s["n_correspondences"] for s in stats] == [s["n_correspondences"] for s in ref_stats], name
            assert rot < 1e-7 and tr < 1e-9, (name, rot, tr)   # arccos(trace) resolves ~2e-8 rad near the identity
        else:
            assert rot < 1e-6 and tr < 1e-6, (name, rot, tr)
        # all ranks hold bit-identical histories (they add the same numbers in the same order)
        mine = torch.from_numpy(hist.copy()).cuda()
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        assert all(bool((g == mine).all()) for g in got), name
        out[name] = dict(outer=len(hist), rot=rot, tr=tr)
    if rank == 0:
        print("MULTI_GPU_OK " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
