"""One-process-per-GPU helpers over the C ABI: the two ways the path partitions.

  * batch of independent scan pairs  -> pairs dealt to ranks, no data-path collective, results gathered at the end
  * one large pair                   -> source points split into contiguous slices, the target replicated; the only
                                        exchange is the 24-moment vector (+ association size) per LM iteration, which the
                                        eval kernel's controller block writes straight into the peers' mailboxes over
                                        NVLink (ppcr_shard_export / ppcr_shard_connect); torch.distributed only carries
                                        the 128-byte mailbox tokens once, at set-up.

torch.distributed is plumbing here (rendezvous, token exchange, result gather); none of it is on the per-iteration path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def slice_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of n items: the first n % world ranks get one extra."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def block_cyclic_indices(n: int, rank: int, world: int, block: int = 4096) -> np.ndarray:
    """Source points of `rank` when the cloud is dealt in blocks of `block` consecutive points, round robin.  A scan is stored
    ring by ring, near rings first, and the cost of a search row follows the local density: contiguous slices give the rank
    that draws the near rings 25 % more search time than the last one (10M-point pair, two GPUs), while a block-cyclic deal
    gives every rank arcs of every ring.  Every rank sorts its own points along the target's Morton curve afterwards, so
    which points a rank holds does not matter for locality as long as they come in runs."""
    n, block = int(n), int(block)
    n_blocks = (n + block - 1) // block
    mine = np.arange(rank, n_blocks, world, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).ravel()
    return idx[idx < n]


def morton_order(cloud: np.ndarray, bits: int = 10) -> np.ndarray:
    """Indices of `cloud` ([N,>=3] float) along a Z-order curve over its bounding cube (`bits` per axis)."""
    xyz = np.asarray(cloud)[:, :3].astype(np.float64)
    lo = xyz.min(axis=0)
    span = float((xyz.max(axis=0) - lo).max()) or 1.0
    q = np.minimum(((xyz - lo) * ((1 << bits) / span)).astype(np.int64), (1 << bits) - 1)
    key = np.zeros(len(xyz), dtype=np.int64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return np.argsort(key, kind="stable")


def morton_chunk_indices(cloud: np.ndarray, rank: int, world: int, chunk: int = 4096) -> np.ndarray:
    """Source points of `rank` when the cloud is cut into runs of `chunk` points ALONG A Z-ORDER CURVE and the runs are dealt
    round robin: every rank gets compact patches at the cloud's full local density (the 32 queries of a warp stay neighbours
    in space, which is what the octree walk's speed depends on) and, there being thousands of patches, an even share of the
    dense and the sparse regions.  Dealing runs of the scan's own storage order (block_cyclic_indices) thins every region by
    the number of ranks instead: 1.85 ms per search and rank on eight GPUs where 1/8 of the single-GPU search is 1.05."""
    order = morton_order(cloud)
    n_chunks = (len(order) + chunk - 1) // chunk
    mine = np.arange(rank, n_chunks, world, dtype=np.int64)
    pos = (mine[:, None] * chunk + np.arange(chunk, dtype=np.int64)[None, :]).ravel()
    return order[pos[pos < len(order)]]


def deal_pairs(n_pairs: int, rank: int, world: int) -> list[int]:
    """Pairs of a batch owned by `rank` (contiguous blocks, same rule as slice_bounds)."""
    lo, hi = slice_bounds(n_pairs, rank, world)
    return list(range(lo, hi))


def gather_tokens(token: bytes, dist=None) -> bytes:
    """All-gathers the mailbox tokens of every rank (rank order).  Works on any backend (gloo / nccl)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    assert len(token) == capi.SHARD_TOKEN_BYTES
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(token), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


class ShardedRegistration(capi.Registration):
    """One rank's share of a single large pair: its slice of the source, the whole target."""

    def __init__(self, source_slice, target, params, rank: int, world: int, options=None, dist=None, **kw):
        super().__init__(source_slice, target, params, options, **kw)
        self.rank, self.world = rank, world
        if world > 1:
            token = (C.c_uint8 * capi.SHARD_TOKEN_BYTES)()
            capi._check(capi.lib().ppcr_shard_export(self._h, rank, world, token))
            tokens = gather_tokens(bytes(token), dist)
            buf = (C.c_uint8 * len(tokens)).from_buffer_copy(tokens)
            capi._check(capi.lib().ppcr_shard_connect(self._h, buf))


def gather_results(local: np.ndarray, dist=None) -> list[np.ndarray]:
    """Gathers per-rank result arrays (e.g. the [n_local,4,4] poses of a batch) on every rank, in rank order."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    objs = [None] * world
    dist.all_gather_object(objs, np.ascontiguousarray(local))
    return objs
