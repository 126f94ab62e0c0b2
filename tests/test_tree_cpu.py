"""The product's octree (csrc/ppcr_tree.h: Morton keys, node split, traversal, top-m lists) compiled for the CPU
and checked against the oracle's FLANN-semantics search.  The GPU kernels call the same functions; only the parallel
drivers around them (sort, per-level launch) differ and those are covered by the -m gpu tests."""
import numpy as np
import pytest

from helpers import emu_tree_search
from probabilistic_point_clouds_registration_b200 import synth


def _check(emu, oracle, src, tgt, radius, m, leaf_cap=32, list_kind=0):
    gi, gd, gc, n_nodes = emu_tree_search(emu, src, tgt, radius, m, leaf_cap, list_kind)
    oi, od, oc, _ = oracle.radius_search(src, tgt, radius, m, use_grid=len(tgt) > 4000)
    assert np.array_equal(gc, oc), np.nonzero(gc != oc)[0][:10]
    w = min(m, oi.shape[1])
    valid = np.arange(w)[None, :] < oc[:, None]
    assert np.array_equal(gi[:, :w][valid], oi[:, :w][valid])
    assert np.array_equal(gd[:, :w][valid].view(np.uint32), od[:, :w][valid].view(np.uint32))
    return gc, n_nodes


@pytest.mark.parametrize("radius,m", [(1.0, 20), (3.0, 20), (0.3, 10), (0.05, 5), (1.0, 1), (1.0, 33), (2.0, 100)])
def test_plane_sphere(emu, oracle, radius, m):
    src, tgt, _ = synth.config1_plane_sphere(seed=1, n_plane=2000, n_sphere=2000)
    _check(emu, oracle, src, tgt, radius, m)


@pytest.mark.parametrize("leaf", [1, 3, 8, 64, 100000])
def test_result_independent_of_tree_shape(emu, oracle, leaf):
    src, tgt, _ = synth.lidar_pair(9, 16, 500)
    _, n_nodes = _check(emu, oracle, src, tgt, 1.0, 12, leaf_cap=leaf)
    assert (n_nodes == 1) == (leaf == 100000)


def test_lidar_with_outliers_and_all_list_kinds(emu, oracle):
    src, tgt, _ = synth.lidar_pair(7, 32, 600, outlier_frac=0.2)
    _check(emu, oracle, src, tgt, 3.0, 20, list_kind=2)
    _check(emu, oracle, src, tgt, 3.0, 20, list_kind=0)
    _check(emu, oracle, src, tgt, 0.5, 10, list_kind=1)
    _check(emu, oracle, src, tgt, 3.0, 50, list_kind=2)
    _check(emu, oracle, src, tgt, 3.0, 20, list_kind=3)
    _check(emu, oracle, src, tgt, 0.5, 10, list_kind=3)
    _check(emu, oracle, src, tgt, 3.0, 1, list_kind=3)
    for kind in (2, 4, 5, 101, 104, 122, 200 + 48, 200 + 100000):
        for m in (1, 2, 3, 7, 10, 16):
            _check(emu, oracle, src[::5], tgt, 2.0, m, list_kind=kind)


@pytest.mark.parametrize("m", [3, 5, 20])
def test_lattice_with_exact_ties(emu, oracle, m):
    g = np.arange(20, dtype=np.float32) * 0.5
    xx, yy = np.meshgrid(g, g, indexing="ij")
    pts = np.ones((400, 4), dtype=np.float32)
    pts[:, 0], pts[:, 1], pts[:, 2] = xx.ravel(), yy.ravel(), 0.0
    _check(emu, oracle, pts, pts, 0.75, m, leaf_cap=4)
    _check(emu, oracle, pts, pts, 0.75, m, leaf_cap=4, list_kind=3)
    _check(emu, oracle, pts, pts, 0.75, m, leaf_cap=4, list_kind=101)
    _check(emu, oracle, pts, pts, 0.75, m, leaf_cap=4, list_kind=112)
    _check(emu, oracle, pts, pts, 0.75, m, leaf_cap=4, list_kind=200 + 100000)
    for kind in (2, 200 + 100000):  # a neighbour at exactly the radius is outside (strict), for the heap walk and the queued phases
        gi, gd, gc, _ = emu_tree_search(emu, pts, pts, 0.5, m, list_kind=kind)
        assert np.all(gc == 1) and np.array_equal(gi[:, 0], np.arange(400))


def test_edge_cases(emu, oracle):
    rng = np.random.default_rng(3)
    tgt = np.ones((50, 4), dtype=np.float32)
    tgt[:, :3] = rng.uniform(-1, 1, (50, 3))
    far = np.ones((7, 4), dtype=np.float32)
    far[:, :3] = rng.uniform(100, 200, (7, 3))
    gi, gd, gc, _ = emu_tree_search(emu, far, tgt, 1.0, 20)
    assert gc.sum() == 0
    small = tgt[:5].copy()
    _check(emu, oracle, tgt, small, 10.0, 20)
    _check(emu, oracle, small[:1], small[:1], 0.5, 3)
    gi, gd, gc, _ = emu_tree_search(emu, tgt, np.zeros((0, 4), dtype=np.float32), 1.0, 4)
    assert gc.sum() == 0
    # many exact duplicates: more points than leaf_cap share one finest cell
    dup = np.concatenate([small] * 9)
    _check(emu, oracle, small, dup, 0.7, 4, leaf_cap=2)
    # large coordinates far from the origin (float fuzz of the binning is covered by the box slack)
    off = tgt.copy()
    off[:, :3] = off[:, :3] * 40 + np.array([5000.0, -3000.0, 800.0], dtype=np.float32)
    _check(emu, oracle, off[::2], off, 6.0, 7, leaf_cap=2)


def test_warm_start_bound_does_not_change_the_result(emu, oracle):
    """The search kernel seeds the pruning bound with (sqrt(previous m-th distance) + displacement)^2.  Any bound
    within which m targets really lie must give the identical result; an infinite one degenerates to the radius."""
    src, tgt, _ = synth.lidar_pair(11, 16, 500)
    m, radius = 8, 1.5
    oi, od, oc, _ = oracle.radius_search(src, tgt, radius, m)
    full = oc == m
    kth = np.where(full, od[:, m - 1], np.inf).astype(np.float32)
    for scale in (1.0, 1.00001, 1.7):
        bounds = np.where(full, kth * np.float32(scale), np.float32(np.inf)).astype(np.float32)
        for kind in (2, 101, 108, 124, 200 + 16, 200 + 100000):
            gi, gd, gc, _ = emu_tree_search(emu, src, tgt, radius, m, bounds=bounds, list_kind=kind)
            assert np.array_equal(gc, oc)
            valid = np.arange(m)[None, :] < oc[:, None]
            assert np.array_equal(gi[valid], oi[valid])


def test_leaf_boxes_never_exceed_a_point_distance(emu, monkeypatch):
    """Every leaf but a root leaf carries the bounding box of its points, and the searches prune a leaf whose box bound exceeds
    the current bound.  That is exact only if the box bound never exceeds the float32 distance of any point inside, for queries
    anywhere: inside the box, next to a face (differences of one ulp), far away, on clouds with coordinates of very different
    magnitude.  emu_tree_search returns -2 when it finds a counter-example."""
    import ctypes as C
    monkeypatch.setenv("EMU_CHECK_LEAF_BOXES", "1")
    rng = np.random.default_rng(7)
    src, tgt, _ = synth.lidar_pair(17, 16, 400)
    # queries displaced from target points along ONE axis: whenever such a point is the extreme of its leaf along that axis the
    # box bound EQUALS its distance, so a bound that is one rounding too large shows up
    base = tgt[rng.choice(len(tgt), 400, replace=False)]
    axial = []
    for axis in range(3):
        for delta in (np.float32(3e-7), np.float32(1e-3), np.float32(0.3)):
            for sign in (-1.0, 1.0):
                q = base.copy()
                q[:, axis] += np.float32(sign) * delta
                axial.append(q)
    far = src[:50].copy()
    far[:, :3] *= 40.0
    queries = np.concatenate([src[:150], far, tgt[:50]] + axial)
    for scale, leaf in ((1.0, 32), (1.0, 4), (1000.0, 8), (1e-3, 8)):
        t = tgt.copy()
        t[:, :3] *= np.float32(scale)
        q = queries.copy()
        q[:, :3] *= np.float32(scale)
        idx = np.full((len(q), 4), -1, dtype=np.int32)
        d2 = np.zeros((len(q), 4), dtype=np.float32)
        cnt = np.zeros(len(q), dtype=np.int32)
        nn = C.c_int(0)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        emu.emu_tree_search.restype = C.c_int64
        rc = emu.emu_tree_search(q.ctypes.data_as(fp), C.c_int64(len(q)), t.ctypes.data_as(fp), C.c_int64(len(t)), C.c_double(0.5 * scale),
                                 C.c_int(4), C.c_int(leaf), C.c_int(2), None, idx.ctypes.data_as(ip), d2.ctypes.data_as(fp),
                                 cnt.ctypes.data_as(ip), C.byref(nn))
        assert rc >= 0, (scale, leaf, rc)
