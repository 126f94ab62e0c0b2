"""The N > 1 host logic on CPU: two gloo ranks (world_size 2, 127.0.0.1).

What is exercised is what the multi-GPU path adds on top of the single-GPU one:
  * the partitioning rules (source slices of a sharded pair, pairs of a batch),
  * the set-up exchange of the mailbox tokens,
  * the algebra the sharded mode rests on: the 24 moments of the weights + normal-equation pass are additive over
    source slices, so rank-ordered sums of per-slice moments reproduce the single-process 7x7 system (the product's
    own eval code, compiled for the CPU, on each rank).
"""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from probabilistic_point_clouds_registration_b200 import multi, synth


def test_slice_bounds_cover_everything_once():
    for n in (0, 1, 7, 1000, 1000064):
        for world in (1, 2, 3, 8):
            spans = [multi.slice_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sum(len(multi.deal_pairs(1024, r, 8)) for r in range(8)) == 1024


def test_block_cyclic_deal_covers_everything_once():
    for n in (0, 1, 100, 10007, 1000064):
        for world in (1, 2, 3, 8):
            for block in (1, 64, 8192):
                parts = [multi.block_cyclic_indices(n, r, world, block) for r in range(world)]
                assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
                assert all(np.all(np.diff(p) > 0) for p in parts if len(p) > 1)
                if n >= world * block * 4:
                    sizes = [len(p) for p in parts]
                    assert max(sizes) - min(sizes) <= block


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import ctypes as C
    import torch
    from helpers import emu_normal_eq
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. token exchange
        from probabilistic_point_clouds_registration_b200 import capi
        token = bytes([rank + 1]) * capi.SHARD_TOKEN_BYTES
        tokens = multi.gather_tokens(token, dist)
        assert tokens == b"".join(bytes([r + 1]) * capi.SHARD_TOKEN_BYTES for r in range(world))
        # 2. additivity of the moments over source slices
        lib = C.CDLL(os.path.join(here, "emu", "libppcr_emu.so"))
        src, tgt, _ = synth.config1_plane_sphere(seed=31, n_plane=900, n_sphere=700)
        data = np.load(os.path.join(out_dir, "assoc.npz"))
        idx, cnt = data["idx"], data["cnt"]
        pose_w = np.array([1.0, 0.01, 0.02, -0.01, 0.01, 0.0, 0.02])
        pose_e = np.array([0.98, 0.03, -0.02, 0.05, 0.03, -0.04, 0.01])
        lo, hi = multi.slice_bounds(len(src), rank, world)
        _, mom = emu_normal_eq(lib, src[lo:hi], tgt, idx[lo:hi], cnt[lo:hi], 5.0, pose_w, pose_e)
        gathered = [torch.zeros(24, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mom))
        total = np.zeros(24)
        for g in gathered:          # rank order, like the mailbox reduction in the controller block
            total += g.numpy()
        _, full = emu_normal_eq(lib, src, tgt, idx, cnt, 5.0, pose_w, pose_e)
        np.testing.assert_allclose(total, full, rtol=1e-12, atol=1e-9)
        # every rank ends with bit-identical sums, hence identical LM decisions
        mine = torch.from_numpy(total.copy())
        other = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        assert all(bool((o == mine).all()) for o in other)
        # 3. batch results come back in rank order
        parts = multi.gather_results(np.full((len(multi.deal_pairs(5, rank, world)), 2), rank), dist)
        assert [len(p) for p in parts] == [3, 2] and all((p == r).all() for r, p in enumerate(parts))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks(emu, oracle, tmp_path):
    src, tgt, _ = synth.config1_plane_sphere(seed=31, n_plane=900, n_sphere=700)
    idx, _, cnt, _ = oracle.radius_search(src, tgt, 1.0, 20)
    np.savez(tmp_path / "assoc.npz", idx=idx, cnt=cnt)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
