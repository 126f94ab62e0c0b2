"""Radius search on the GPU grid vs the oracle (FLANN semantics): neighbour index sets must be
bit-exact, rows sorted by (d2, index), distances bit-identical float32."""
import numpy as np
import pytest

from helpers import rows_as_sets
from probabilistic_point_clouds_registration_b200 import synth

pytestmark = pytest.mark.gpu


def _compare(capi, oracle, src, tgt, radius, m, leaf_capacity=0):
    gi, gd, gc = capi.radius_search(src, tgt, radius, m, leaf_capacity)
    oi, od, oc, _ = oracle.radius_search(src, tgt, radius, m, use_grid=len(tgt) > 4000)
    cap = oi.shape[1]
    assert np.array_equal(gc, oc), f"counts differ at {np.nonzero(gc != oc)[0][:10]}"
    k = min(m, cap)
    for i in np.nonzero(oc > 0)[0]:
        c = oc[i]
        assert np.array_equal(gi[i, :c], oi[i, :c]), (i, gi[i, :c], oi[i, :c])
        assert np.array_equal(gd[i, :c].view(np.uint32), od[i, :c].view(np.uint32)), i
    assert np.all(gi[:, :k][np.arange(k)[None, :] >= gc[:, None]] == -1)
    return gc


@pytest.mark.parametrize("radius,m", [(1.0, 20), (3.0, 20), (0.3, 10), (0.05, 5), (1.0, 1), (1.0, 32), (1.0, 33),
                                      (1.0, 48), (1.0, 64), (2.0, 100)])
def test_plane_sphere(capi, oracle, radius, m):
    src, tgt, _ = synth.config1_plane_sphere(seed=1, n_plane=3000, n_sphere=3000)
    cnt = _compare(capi, oracle, src, tgt, radius, m)
    if radius >= 1.0 and m <= 64:
        assert (cnt == m).mean() > 0.9  # saturated regime


@pytest.mark.parametrize("outliers", [0.0, 0.2])
def test_lidar_like(capi, oracle, outliers):
    src, tgt, _ = synth.lidar_pair(7, 32, 600, outlier_frac=outliers)
    _compare(capi, oracle, src, tgt, 3.0, 20)
    _compare(capi, oracle, src, tgt, 0.5, 10)


@pytest.mark.parametrize("leaf", [1, 3, 8, 64, 100000])
def test_result_independent_of_tree_shape(capi, oracle, leaf):
    src, tgt, _ = synth.lidar_pair(9, 16, 500)
    _compare(capi, oracle, src, tgt, 1.0, 12, leaf_capacity=leaf)


@pytest.mark.parametrize("m", [3, 5, 20])
def test_lattice_with_exact_ties(capi, oracle, m):
    # a flat 0.5-spaced lattice (the spacing of the reference's own fixture): every point has 4 neighbours at
    # exactly d2 = 0.25 and 4 at exactly 0.5.  Ties inside the kept set are ordered by index and a tie at the
    # m-th boundary keeps the lower index; the radius bound is strict (0.75^2 excludes nothing at 0.5, r = 0.5
    # excludes the four axis neighbours because 0.25 < 0.25 is false).
    g = np.arange(20, dtype=np.float32) * 0.5
    xx, yy = np.meshgrid(g, g, indexing="ij")
    pts = np.ones((400, 4), dtype=np.float32)
    pts[:, 0], pts[:, 1], pts[:, 2] = xx.ravel(), yy.ravel(), 0.0
    _compare(capi, oracle, pts, pts, 0.75, m)
    gi, gd, gc = capi.radius_search(pts, pts, 0.5, m)
    assert np.all(gc == 1) and np.array_equal(gi[:, 0], np.arange(400))


def test_large_lattice_ties_on_the_oracle_grid_path(capi, oracle):
    """The same kind of lattice above 4000 points, where the oracle answers from its CPU grid instead of brute force:
    both sides keep the pure (d2, index) order, so a tie at the m-th boundary resolves to the lower index whatever the
    traversal order (a 3-D lattice: 6 neighbours at d2 = 0.25, 12 at 0.5, 8 at 0.75)."""
    g = np.arange(18, dtype=np.float32) * 0.5
    xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
    pts = np.ones((18 ** 3, 4), dtype=np.float32)
    pts[:, 0], pts[:, 1], pts[:, 2] = xx.ravel(), yy.ravel(), zz.ravel()
    assert len(pts) > 4000
    for radius, m in ((0.75, 4), (0.75, 10), (0.9, 20), (0.9, 32)):
        _compare(capi, oracle, pts, pts, radius, m)
    # and shuffled, so that index order and Morton order disagree
    perm = np.random.default_rng(5).permutation(len(pts))
    _compare(capi, oracle, pts[::3], pts[perm], 0.75, 10)


def test_edge_cases(capi, oracle):
    rng = np.random.default_rng(3)
    tgt = np.ones((50, 4), dtype=np.float32)
    tgt[:, :3] = rng.uniform(-1, 1, (50, 3))
    far = np.ones((7, 4), dtype=np.float32)
    far[:, :3] = rng.uniform(100, 200, (7, 3))
    gi, gd, gc = capi.radius_search(far, tgt, 1.0, 20)
    assert gc.sum() == 0 and np.all(gi == -1)
    # fewer targets than max_neighbours: capacity becomes n_tgt (pcl: max_nn > N_t => all points)
    small = tgt[:5].copy()
    _compare(capi, oracle, tgt, small, 10.0, 20)
    # one target, one source, coincident
    _compare(capi, oracle, small[:1], small[:1], 0.5, 3)
    # empty clouds
    gi, gd, gc = capi.radius_search(tgt, np.zeros((0, 4), dtype=np.float32), 1.0, 4)
    assert gc.sum() == 0
    gi, gd, gc = capi.radius_search(np.zeros((0, 4), dtype=np.float32), tgt, 1.0, 4)
    assert len(gc) == 0
    # duplicates in the target: ties broken by index
    dup = np.concatenate([small, small, small])
    _compare(capi, oracle, small, dup, 0.7, 4)


def test_unsupported_max_neighbours(capi):
    tgt = np.ones((10, 4), dtype=np.float32)
    for bad in (0, -1, 129):
        with pytest.raises(capi.PpcrError) as e:
            capi.radius_search(tgt, tgt, 1.0, bad)
        assert e.value.code == 4


def test_full_size_properties(capi, oracle):
    """BASELINE config 3 scale (1M points): size-independent properties + a sampled oracle check."""
    src, tgt, _ = synth.config3_lidar_1m()
    radius, m = 0.5, 10
    gi, gd, gc = capi.radius_search(src, tgt, radius, m)
    r2 = np.float32(radius * radius)
    valid = np.arange(m)[None, :] < gc[:, None]
    assert gc.min() >= 0 and gc.max() <= m
    assert np.all(gd[valid] < r2)                                   # strict radius bound
    d = np.where(valid, gd, np.inf)
    assert np.all(np.diff(d, axis=1)[valid[:, 1:]] >= 0)            # rows sorted by distance
    assert np.all(gi[valid] >= 0) and np.all(gi[valid] < len(tgt))
    # recomputed distances are the float32 no-FMA distances of the returned indices
    rows = np.repeat(np.arange(len(src)), m).reshape(len(src), m)[valid]
    a = src[rows, :3]
    b = tgt[gi[valid], :3]
    dd = a - b
    d2 = (dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]
    assert np.array_equal(d2.astype(np.float32).view(np.uint32), gd[valid].view(np.uint32))
    # sampled exact check against the oracle
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(len(src), 20000, replace=False))
    oi, od, oc, _ = oracle.radius_search(src[sel], tgt, radius, m, use_grid=True)
    assert np.array_equal(oc, gc[sel])
    assert rows_as_sets(oi, oc) == rows_as_sets(gi[sel], gc[sel])


def test_ten_million_point_pair_sampled_rows(capi, oracle):
    """BASELINE config 4's clouds (10M points each, -m 10 -r 0.5): size-independent properties of every row and an
    exact comparison of 4000 sampled rows with the oracle's grid search over the whole 10M-point target."""
    src, tgt, _ = synth.config4_lidar_10m()
    radius, m = 0.5, 10
    gi, gd, gc = capi.radius_search(src, tgt, radius, m)
    r2 = np.float32(radius * radius)
    valid = np.arange(m)[None, :] < gc[:, None]
    assert gc.min() >= 0 and gc.max() <= m
    assert np.all(gd[valid] < r2)
    d = np.where(valid, gd, np.inf)
    assert np.all(np.diff(d, axis=1)[valid[:, 1:]] >= 0)
    assert np.all(gi[valid] >= 0) and np.all(gi[valid] < len(tgt))
    assert np.all(gi[~valid] == -1)
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(len(src), 4000, replace=False))
    oi, od, oc, _ = oracle.radius_search(src[sel], tgt, radius, m, use_grid=True)
    assert np.array_equal(oc, gc[sel])
    for k, i in enumerate(sel):
        c = oc[k]
        assert np.array_equal(gi[i, :c], oi[k, :c]), i
        assert np.array_equal(gd[i, :c].view(np.uint32), od[k, :c].view(np.uint32)), i


def _moved_search_rows(capi, src, tgt, n_iter, **kw):
    """Association of the n_iter-th search of an align() and the (moved) cloud that search saw."""
    with capi.Registration(src, tgt, capi.make_params(n_iter=n_iter - 1, **kw)) as reg:
        reg.align()
        cloud = reg.filtered_source()  # after n_iter - 1 increments: what search n_iter starts from
    with capi.Registration(src, tgt, capi.make_params(n_iter=n_iter, **kw)) as reg:
        reg.align()
        idx, cnt = reg.association()
        n_done = len(reg.iteration_stats())
    return cloud, idx, cnt, n_done


# the queued kernel's three ways through a chunk: queues (default), every chunk walked with heaps (threshold 0), and
# queues with a candidate list so short that most queries overflow it and fall back one by one
QUEUED_MODES = {"off": {"PPCR_SEARCH_QUEUED": "0"}, "queues": {"PPCR_SEARCH_QUEUED": "1"},
                "all_heavy": {"PPCR_SEARCH_QUEUED": "1", "PPCR_Q_HEAVY": "0"},
                "overflow": {"PPCR_SEARCH_QUEUED": "1", "PPCR_Q_HEAVY": "1e9", "PPCR_Q_CAND": "12"}}


@pytest.mark.parametrize("mode", sorted(QUEUED_MODES))
@pytest.mark.parametrize("radius,m", [(0.5, 10), (1.5, 20), (3.0, 5)])
def test_searches_after_a_cloud_move(capi, oracle, monkeypatch, mode, radius, m):
    """Every search of an align() but the first is fused with the cloud move and pruned by a bound taken from the previous
    association (the farthest previous neighbour of the moved query).  Its rows must still be exactly the m nearest
    in-radius targets of the moved cloud -- for k_search and for every path through the queued kernel."""
    for k, v in QUEUED_MODES[mode].items():
        monkeypatch.setenv(k, v)
    src, tgt, _ = synth.lidar_pair(23, 32, 700, yaw_deg=1.0, trans=(0.2, 0.05, 0.0))
    for n_iter in (2, 4):
        cloud, idx, cnt, n_done = _moved_search_rows(capi, src, tgt, n_iter, max_neighbours=m, radius=radius, dof=5.0)
        assert n_done == n_iter
        oi, od, oc, _ = oracle.radius_search(cloud, tgt, radius, m, use_grid=True)
        assert np.array_equal(cnt, oc)
        w = min(m, oi.shape[1], idx.shape[1])
        valid = np.arange(w)[None, :] < oc[:, None]
        got = np.sort(np.where(valid, idx[:, :w], -1), axis=1)
        want = np.sort(np.where(valid, oi[:, :w], -1), axis=1)
        assert np.array_equal(got, want)
