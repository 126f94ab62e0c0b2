"""Work counters of the queued search (phases A/B/C of k_search_q, CPU build of csrc/ppcr_tree.h) on queries that MOVED since
their last search: the bound is the distance of the farthest previous neighbour to the moved query, as in the kernel.

    python tools/tree_stats_moved.py [c3|c5] [n_queries] [move_m] [leaf_cap ...]
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

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
move = float(sys.argv[3]) if len(sys.argv) > 3 else 0.017
leaf_caps = [int(a) for a in sys.argv[4:]] or [32]
extra = os.environ.get("EMU_FLAGS", "").split()
so = "/tmp/libppcr_emu_stats.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DPPCR_TREE_STATS", *extra, "-o", so,
                       os.path.join(ROOT, "tests", "emu", "emu_host_logic.cpp")])
lib = C.CDLL(so)
src, tgt = bench.make_pair(workload, 0)
prm = bench.WORKLOADS[workload]["params"]
m = prm["max_neighbours"]
rng = np.random.default_rng(0)
sel = rng.choice(len(src), size=min(nq, len(src)), replace=False)
# queries close to the target surface (as after a few outer iterations): target points + small noise, then the move
q0 = tgt[rng.choice(len(tgt), size=len(sel), replace=False)].copy()
q0[:, :3] += rng.normal(scale=0.01, size=(len(q0), 3)).astype(np.float32)
names = ["opens", "leaves", "leaves_skipped", "points", "survivors", "inserts", "stack_skipped"]


def stats():
    out = (C.c_longlong * 7)()
    lib.emu_tree_stats(out, 1)
    return np.array(list(out), dtype=np.float64)


for leaf in leaf_caps:
    idx, d2, cnt, n_nodes = helpers.emu_tree_search(lib, q0, tgt, prm["radius"], m, leaf_cap=leaf)
    stats()
    d = np.array([0.6, 0.7, 0.3])
    q1 = q0.copy()
    q1[:, :3] += (move * d / np.linalg.norm(d)).astype(np.float32)
    full = cnt == m
    nb = tgt[np.where(idx >= 0, idx, 0)][:, :, :3]
    dd = ((q1[:, None, :3] - nb) ** 2).sum(axis=2)
    bound = np.where(full, dd.max(axis=1), np.float32(prm["radius"] ** 2)).astype(np.float32) * np.float32(1.00001)
    helpers.emu_tree_search(lib, q1, tgt, prm["radius"], m, leaf_cap=leaf, list_kind=264, bounds=bound)
    s = stats() / len(q1)
    print(f"{workload} leaf_cap={leaf} move={move} nodes={n_nodes}: " + "  ".join(f"{n}={x:.1f}" for n, x in zip(names, s)))
