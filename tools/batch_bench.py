"""Throughput of a batch of independent c5 pairs (BASELINE configs[4]) through ppcr_align_batch on one GPU.

    python tools/batch_bench.py [n_pairs] [slots ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from probabilistic_point_clouds_registration_b200 import capi, synth  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
slot_list = [int(a) for a in sys.argv[2:]] or [1, 4, 8, 16]
t0 = time.perf_counter()
pairs = []
for i in range(n_pairs):
    s, t, _ = synth.config5_pair(i)
    pairs.append((np.ascontiguousarray(s), np.ascontiguousarray(t)))
print(f"generated {n_pairs} pairs of {len(pairs[0][0])} points in {time.perf_counter() - t0:.1f} s")
params = capi.make_params(**bench.WORKLOADS["c5"]["params"])
# warm-up: every lane's share of the memory pool exists, and the clocks are up (the GPU idled while the pairs were generated:
# a timed run right after a short warm-up came out anywhere between 150 and 450 pairs/s)
t0 = time.perf_counter()
while time.perf_counter() - t0 < 2.0:
    capi.align_batch(pairs[:min(n_pairs, 8 * max(slot_list))], params, slots=max(slot_list))
for slots in slot_list:
    t0 = time.perf_counter()
    T, n_outer, corr = capi.align_batch(pairs, params, slots=slots)
    dt = time.perf_counter() - t0
    print(f"slots={slots:3d}: {dt * 1e3:8.1f} ms for {n_pairs} pairs -> {n_pairs / dt:7.1f} pairs/s, {corr.sum() / dt / 1e9:6.2f} G corr/s, "
          f"outer iterations {n_outer.min()}..{n_outer.max()}")
