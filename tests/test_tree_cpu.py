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
