"""Shared helpers for the test-suite."""
import ctypes as C

import numpy as np


class EmuStats(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("cost_drop", C.c_double),
                ("n_correspondences", C.c_int64), ("lm_iterations", C.c_int32), ("num_successful_steps", C.c_int32)]


def emu_align(lib, src, tgt, m, dof, radius, n_iter=1000, thr=0.01, nd=5.0, ftol=1e-5, fast=0,
              x0=(1, 0, 0, 0, 0, 0, 0)):
    src = np.ascontiguousarray(src, dtype=np.float32)
    tgt = np.ascontiguousarray(tgt, dtype=np.float32)
    cap = max(1, n_iter)
    hist = np.zeros((cap, 16))
    st = (EmuStats * cap)()
    out = np.zeros_like(src)
    x0 = np.array(x0, dtype=np.float64)
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    lib.emu_align.restype = C.c_int
    n = lib.emu_align(src.ctypes.data_as(fp), C.c_int64(len(src)), tgt.ctypes.data_as(fp), C.c_int64(len(tgt)),
                      C.c_int(m), C.c_double(dof), C.c_double(radius), C.c_int(n_iter), C.c_double(thr),
                      C.c_double(nd), x0.ctypes.data_as(dp), C.c_double(ftol), C.c_int(fast),
                      hist.ctypes.data_as(dp), st, C.c_int(cap), out.ctypes.data_as(fp))
    k = min(n, cap)
    stats = [dict(initial_cost=s.initial_cost, final_cost=s.final_cost, cost_drop=s.cost_drop,
                  n_correspondences=s.n_correspondences, lm_iterations=s.lm_iterations,
                  num_successful_steps=s.num_successful_steps) for s in st[:k]]
    return n, hist[:k].reshape(k, 4, 4), stats, out


def emu_normal_eq(lib, src, tgt, idx, cnt, dof, pose_w, pose_e, fast=0):
    src = np.ascontiguousarray(src, dtype=np.float32)
    tgt = np.ascontiguousarray(tgt, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    pw = np.ascontiguousarray(pose_w, dtype=np.float64)
    pe = np.ascontiguousarray(pose_e, dtype=np.float64)
    ne = np.zeros(36)
    mom = np.zeros(24)
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.emu_normal_eq.restype = None
    lib.emu_normal_eq(src.ctypes.data_as(fp), C.c_int64(len(src)), tgt.ctypes.data_as(fp), C.c_int64(len(tgt)),
                      idx.ctypes.data_as(ip), cnt.ctypes.data_as(ip), C.c_int(idx.shape[1]), C.c_double(dof),
                      pw.ctypes.data_as(dp), pe.ctypes.data_as(dp), C.c_int(fast), ne.ctypes.data_as(dp),
                      mom.ctypes.data_as(dp))
    return ne, mom


def csr_from_rows(idx, cnt):
    """[n,m] padded rows + counts -> (row_ptr int64, col int32) with each row sorted by column, like Eigen's
    setFromTriplets leaves it (src/prob_point_cloud_registration.cc:82-83)."""
    row_ptr = np.zeros(len(cnt) + 1, dtype=np.int64)
    row_ptr[1:] = np.cumsum(cnt)
    col = np.zeros(int(row_ptr[-1]), dtype=np.int32)
    for i, c in enumerate(cnt):
        col[row_ptr[i]:row_ptr[i + 1]] = np.sort(idx[i, :c])
    return row_ptr, col


def pose_delta(Ta, Tb):
    """(rotation angle in rad, translation distance in m) between two 4x4 transforms."""
    dR = Ta[:3, :3].T @ Tb[:3, :3]
    c = np.clip((np.trace(dR) - 1.0) / 2.0, -1.0, 1.0)
    return float(np.arccos(c)), float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))


def rows_as_sets(idx, cnt):
    return [frozenset(idx[i, :cnt[i]].tolist()) for i in range(len(cnt))]


def emu_tree_search(lib, src, tgt, radius, m, leaf_cap=32, list_kind=2, bounds=None):
    """The product's octree build + traversal (csrc/ppcr_tree.h) compiled for the CPU.  bounds: optional per-query
    squared distance within which m targets are known to lie (the warm start of the search kernel)."""
    src = np.ascontiguousarray(src, dtype=np.float32)
    tgt = np.ascontiguousarray(tgt, dtype=np.float32)
    idx = np.full((len(src), m), -1, dtype=np.int32)
    d2 = np.zeros((len(src), m), dtype=np.float32)
    cnt = np.zeros(len(src), dtype=np.int32)
    n_nodes = C.c_int(0)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    lib.emu_tree_search.restype = C.c_int64
    lib.emu_tree_search(src.ctypes.data_as(fp), C.c_int64(len(src)), tgt.ctypes.data_as(fp), C.c_int64(len(tgt)),
                        C.c_double(radius), C.c_int(m), C.c_int(leaf_cap), C.c_int(list_kind),
                        None if bounds is None else np.ascontiguousarray(bounds, dtype=np.float32).ctypes.data_as(fp),
                        idx.ctypes.data_as(ip),
                        d2.ctypes.data_as(fp), cnt.ctypes.data_as(ip), C.byref(n_nodes))
    return idx, d2, cnt, n_nodes.value
