"""Work counters of the queued search on the REAL in-loop state of a bench pair: the source as the reference leaves it after
`k` outer iterations (CPU oracle), its association, and the increment of iteration k + 1.  Prints per query: nodes opened,
leaf tasks, points tested, candidates; and the distribution of candidates per query.

    python tools/tree_stats_real.py [c3|c5] [k] [n_queries] [leaf_cap]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import helpers  # noqa: E402
from oracle import oracle as O  # noqa: E402
from probabilistic_point_clouds_registration_b200 import synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 40000
leaf = int(sys.argv[4]) if len(sys.argv) > 4 else 32
extra = os.environ.get("EMU_FLAGS", "").split()
so = "/tmp/libppcr_emu_stats.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DPPCR_TREE_STATS", *extra, "-o", so,
                       os.path.join(ROOT, "tests", "emu", "emu_host_logic.cpp")])
lib = C.CDLL(so)
src, tgt = bench.make_pair(workload, 0)
prm = bench.WORKLOADS[workload]["params"]
m = prm["max_neighbours"]
cache = f"/tmp/real_{workload}_{k}.npz"
if os.path.exists(cache):
    z = np.load(cache)
    q0, dT = z["q0"], z["dT"]
else:
    r = O.align(src, tgt, O.make_params(n_iter=k + 1, **prm), O.make_options(inner_kind=1), use_grid=True)
    rk = O.align(src, tgt, O.make_params(n_iter=k, **prm), O.make_options(inner_kind=1), use_grid=True)
    q0 = rk.filtered_source
    dT = r.history[k] @ np.linalg.inv(r.history[k - 1])
    np.savez(cache, q0=q0, dT=dT)
print("increment translation (m):", np.linalg.norm(dT[:3, 3]), "rotation (deg):",
      np.degrees(np.arccos(np.clip((np.trace(dT[:3, :3]) - 1) / 2, -1, 1))))
rng = np.random.default_rng(0)
# a contiguous run of the Morton-ish order is not needed for counters: a random sample of the real queries
sel = rng.choice(len(q0), size=min(nq, len(q0)), replace=False)
q0 = np.ascontiguousarray(q0[sel])
q1 = synth.apply_T_like_pcl(q0, dT)
print("mean query displacement (m):", np.linalg.norm(q1[:, :3] - q0[:, :3], axis=1).mean())
names = ["opens", "leaves", "leaves_skipped", "points", "survivors", "inserts", "stack_skipped"]


def stats():
    out = (C.c_longlong * 7)()
    lib.emu_tree_stats(out, 1)
    return np.array(list(out), dtype=np.float64)


idx, d2, cnt, n_nodes = helpers.emu_tree_search(lib, q0, tgt, prm["radius"], m, leaf_cap=leaf)
stats()
full = cnt == m
nb = tgt[np.where(idx >= 0, idx, 0)][:, :, :3]
dd = ((q1[:, None, :3].astype(np.float32) - nb) ** 2).sum(axis=2)
bound = np.where(full, dd.max(axis=1), np.float32(prm["radius"] ** 2)).astype(np.float32) * np.float32(1.00001)
kth_old = np.where(full, d2[:, m - 1], np.inf)
idx1, d21, cnt1, _ = helpers.emu_tree_search(lib, q1, tgt, prm["radius"], m, leaf_cap=leaf, list_kind=264, bounds=bound)
s = stats() / len(q1)
print(f"{workload} after {k} iterations, leaf_cap={leaf}: saturated rows {full.mean():.3f}; " + "  ".join(f"{n}={x:.1f}" for n, x in zip(names, s)))
op = np.zeros(len(q1), dtype=np.int32)
lv = np.zeros(len(q1), dtype=np.int32)
lib.emu_tree_per_query(op.ctypes.data_as(C.POINTER(C.c_int)), lv.ctypes.data_as(C.POINTER(C.c_int)), C.c_longlong(len(q1)))
for name, v in (("opens", op), ("leaf tasks", lv)):
    qs = [0.5, 0.75, 0.9, 0.95, 0.98, 0.99, 0.999]
    print(f"per-query {name}: mean {v.mean():.1f}  " + "  ".join(f"p{int(1000*q)/10:g}={np.quantile(v, q):.0f}" for q in qs) + f"  max={v.max()}")
for cap in (4, 6, 8, 12, 16):
    print(f"  queries with more than {cap} opens: {np.mean(op > cap):.3f}; share of all opens beyond the cap: {np.maximum(op - cap, 0).sum() / op.sum():.3f}; "
          f"P(a warp of 32 holds one) ~ {1 - (1 - np.mean(op > cap)) ** 32:.2f}")
kth_new = np.where(cnt1 == m, d21[:, m - 1], np.inf)
ok = full & (cnt1 == m)
ratio = bound[ok] / kth_new[ok]
print("bound / new m-th distance (squared): mean %.2f median %.2f p90 %.2f p99 %.2f" % (ratio.mean(), np.median(ratio), np.quantile(ratio, 0.9), np.quantile(ratio, 0.99)))
print("unsaturated rows (bound = r^2):", (~full).mean())
# ideal: the bound is the new m-th distance itself
helpers.emu_tree_search(lib, q1, tgt, prm["radius"], m, leaf_cap=leaf, list_kind=264, bounds=np.where(np.isfinite(kth_new), kth_new, np.float32(prm["radius"] ** 2)).astype(np.float32))
s = stats() / len(q1)
print("with the exact bound:     " + "  ".join(f"{n}={x:.1f}" for n, x in zip(names, s)))
# split: saturated vs unsaturated rows
for label, mask in (("saturated", full), ("unsaturated", ~full)):
    if mask.sum() == 0:
        continue
    helpers.emu_tree_search(lib, np.ascontiguousarray(q1[mask]), tgt, prm["radius"], m, leaf_cap=leaf, list_kind=264,
                            bounds=np.ascontiguousarray(bound[mask]))
    s = stats() / mask.sum()
    print(f"{label:12s} ({mask.mean():.3f} of rows): " + "  ".join(f"{n}={x:.1f}" for n, x in zip(names, s)))
