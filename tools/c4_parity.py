"""Association of the k-th search of the 10M-point pair (BASELINE configs[3]) from both search kernels against the oracle's grid
search on the cloud that search saw (one-off check at full size; tests/test_gpu_search.py does the same on small clouds).

    python tools/c4_parity.py [k] [rings] [az]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from probabilistic_point_clouds_registration_b200 import capi, synth  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rings = int(sys.argv[2]) if len(sys.argv) > 2 else 320
az = int(sys.argv[3]) if len(sys.argv) > 3 else 31250
src, tgt, _ = synth.lidar_pair(4, rings, az)
kw = dict(max_neighbours=10, radius=0.5, dof=5.0)
for mode in ("1", "0"):
    os.environ["PPCR_SEARCH_QUEUED"] = mode
    # the cloud search k saw IN THIS MODE (the two kernels store a row's neighbours in different orders, so the float32 row sums,
    # the poses and hence the moved clouds of the two modes differ in the last bits)
    with capi.Registration(src, tgt, capi.make_params(n_iter=k - 1, **kw)) as reg:
        reg.align()
        cloud = reg.filtered_source()
    oi, od, oc, _ = O.radius_search(cloud, tgt, 0.5, 10, use_grid=True)
    print(f"oracle: K = {int(oc.sum())}", flush=True)
    with capi.Registration(src, tgt, capi.make_params(n_iter=k, **kw)) as reg:
        reg.align()
        idx, cnt = reg.association()
        cloud2 = reg.filtered_source()
    bad = np.nonzero(cnt != oc)[0]
    w = min(10, oi.shape[1], idx.shape[1])
    valid = np.arange(w)[None, :] < oc[:, None]
    got = np.sort(np.where(valid, idx[:, :w], -1), axis=1)
    want = np.sort(np.where(valid, oi[:, :w], -1), axis=1)
    rows = np.nonzero((got != want).any(axis=1))[0]
    print(f"PPCR_SEARCH_QUEUED={mode}: K = {int(cnt.sum())}, rows with another count {len(bad)}, rows with another set {len(rows)}", flush=True)
    for i in rows[:5]:
        print("   row", i, "q", cloud[i], "count", cnt[i], oc[i], "\n      got ", idx[i], "\n      want", oi[i], "\n      d2  ", od[i])
