"""Work counters of the octree traversal on a bench workload (CPU build of csrc/ppcr_tree.h, no GPU needed).

    python tools/tree_stats.py [c3|c1|c5] [n_queries] [leaf_cap ...]

Prints, per query: nodes opened, leaves scanned, points tested, survivors of the first pass, list insertions -- for a
cold search (bound = r^2) and for a warm one (bound = the true m-th distance, what a converged iteration sees).
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
leaf_caps = [int(a) for a in sys.argv[3:]] or [32]
so = "/tmp/libppcr_emu_stats.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DPPCR_TREE_STATS", "-o", so,
                       os.path.join(ROOT, "tests", "emu", "emu_host_logic.cpp")])
lib = C.CDLL(so)
src, tgt = bench.make_pair(workload, 0)
prm = bench.WORKLOADS[workload]["params"]
rng = np.random.default_rng(0)
start = int(rng.integers(0, max(1, len(src) - nq)))
# a contiguous block of the Morton-sorted source would be ideal; a random sample is representative of the averages
sel = rng.choice(len(src), size=min(nq, len(src)), replace=False)
q = src[sel]
names = ["opens", "leaves", "leaves_skipped", "points", "survivors", "inserts", "stack_skipped"]


def stats():
    out = (C.c_longlong * 7)()
    lib.emu_tree_stats(out, 1)
    return np.array(list(out), dtype=np.float64)


for leaf in leaf_caps:
    idx, d2, cnt, n_nodes = helpers.emu_tree_search(lib, q, tgt, prm["radius"], prm["max_neighbours"], leaf_cap=leaf)
    cold = stats() / len(q)
    m = prm["max_neighbours"]
    kth = np.where(cnt == m, d2[np.arange(len(q)), np.maximum(cnt - 1, 0)], np.float32(prm["radius"] ** 2))
    helpers.emu_tree_search(lib, q, tgt, prm["radius"], m, leaf_cap=leaf, bounds=kth.astype(np.float32))
    warm = stats() / len(q)
    print(f"{workload} leaf_cap={leaf} nodes={n_nodes} mean cnt={cnt.mean():.2f}")
    for label, v in (("cold", cold), ("warm", warm)):
        print("  " + label + "  " + "  ".join(f"{n}={x:.1f}" for n, x in zip(names, v)))
