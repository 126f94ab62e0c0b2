"""The oracle against every golden vector / known-answer test the reference holds for the hot path
(G1-G4) plus self-consistency of its own parts.  CPU only."""
import numpy as np
import pytest

from helpers import csr_from_rows
from probabilistic_point_clouds_registration_b200 import synth

ROW_PTR = np.array([0, 3, 7])
SQ_ERR = np.array([1, 1, 1, 1, 4, 9, 16.0])   # test/ProbabilisticWeightsTest.cc:16


def test_g1_t_distribution_weights(oracle):
    w = oracle.update_weights(ROW_PTR, SQ_ERR, 5.0, 1)          # ProbabilisticWeights(5, 1, 4), T_W:39
    np.testing.assert_allclose(w[:3], [1 / 3] * 3, atol=1e-6)
    np.testing.assert_allclose(w[3:], [0.7151351, 0.1412613, 0.0241258, 0.0047656], atol=1e-6)   # T_W:42-43


def test_g2_gaussian_weights(oracle):
    w = oracle.update_weights(ROW_PTR, SQ_ERR, np.inf, 1)       # T_W:55
    np.testing.assert_allclose(w[:3], [1 / 3] * 3, atol=1e-6)
    np.testing.assert_allclose(w[3:], [0.805153702921689, 0.179654074677018, 0.0147469044726408,
                                       0.000445317928652638], atol=1e-6)                         # T_W:59-60


@pytest.mark.parametrize("dof,dim", [(5.0, 1), (5.0, 3), (0.7, 3), (np.inf, 3)])
def test_weights_closed_form(oracle, dof, dim):
    rng = np.random.default_rng(0)
    se = rng.uniform(0, 9, 11)
    w = oracle.update_weights(np.array([0, 11]), se, dof, dim)
    np.testing.assert_allclose(w, oracle.weights_closed_form(se, dof, dim), rtol=1e-12)
    if np.isinf(dof):
        assert abs(w.sum() - 1) < 1e-12       # Gaussian rows are a softmax; t rows are not


@pytest.mark.parametrize("dof", [np.inf, 5.0])
@pytest.mark.parametrize("kind", [0, 1])
def test_g3_g4_exact_association_fixture(oracle, dof, kind):
    """test/PointCloudRegistrationTest.cc:30-116 restated: mean point error < 1e-6 after the solve."""
    src = synth.reference_test_cloud()
    T = synth.reference_test_transform()
    np.testing.assert_allclose(T[:3, 3], [2.35688666, 0.83371773, 0.0], atol=1e-8)
    tgt = synth.apply_T_like_pcl(src, T)
    n = len(src)
    r = oracle.iteration_solve(src, tgt, np.arange(n + 1), np.arange(n, dtype=np.int32),
                               oracle.make_params(max_neighbours=3, dof=dof),
                               oracle.make_options(function_tolerance=10e-5, inner_kind=kind))
    aligned = oracle.transform(src, r.T)
    err = np.sqrt(((tgt[:, :3].astype(np.float64) - aligned[:, :3].astype(np.float64)) ** 2).sum(1)).mean()
    assert err < 1e-6
    assert synth.pose_error(r.T, T)[1] < 1e-7


def test_inner_solvers_agree(oracle):
    """dual-number + dense QR (faithful) and analytic + normal equations (fast) take the same LM path."""
    src, tgt, _ = synth.config1_plane_sphere(seed=3, n_plane=500, n_sphere=400)
    idx, _, cnt, _ = oracle.radius_search(src, tgt, 1.0, 20)
    row_ptr, col = csr_from_rows(idx, cnt)
    p = oracle.make_params()
    a = oracle.iteration_solve(src, tgt, row_ptr, col, p, oracle.make_options(inner_kind=0))
    b = oracle.iteration_solve(src, tgt, row_ptr, col, p, oracle.make_options(inner_kind=1))
    assert a.num_iterations == b.num_iterations and a.termination == b.termination
    np.testing.assert_allclose(a.T, b.T, atol=1e-10)
    np.testing.assert_allclose([a.initial_cost, a.final_cost], [b.initial_cost, b.final_cost], rtol=1e-10)


@pytest.mark.parametrize("radius,m", [(1.0, 20), (0.3, 7), (3.0, 20), (1.0, 1)])
def test_grid_search_equals_brute_force(oracle, radius, m):
    src, tgt, _ = synth.config1_plane_sphere(seed=2, n_plane=1200, n_sphere=800)
    bi, bd, bc, bt = oracle.radius_search(src, tgt, radius, m, use_grid=False)
    gi, gd, gc, gt = oracle.radius_search(src, tgt, radius, m, use_grid=True)
    assert bt == gt and np.array_equal(bc, gc) and np.array_equal(bi, gi) and np.array_equal(bd, gd)
    r2 = np.float32(radius * radius)
    valid = np.arange(bi.shape[1])[None, :] < bc[:, None]
    assert np.all(bd[valid] < r2)
    # brute-force numpy check of one row: strict radius, (d2, idx) order, float32 no-FMA distance
    i = 17
    d = src[i, :3] - tgt[:, :3]
    d2 = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
    order = np.lexsort((np.arange(len(tgt)), d2))
    order = order[d2[order] < r2][:m]
    assert np.array_equal(bi[i, :bc[i]], order)


def test_unlimited_max_nn(oracle):
    """max_nn == 0 or > N_t means "all in-radius points" (pcl::KdTreeFLANN::radiusSearch)."""
    src, tgt, _ = synth.config1_plane_sphere(seed=2, n_plane=60, n_sphere=40)
    a = oracle.radius_search(src, tgt, 3.0, 0)
    b = oracle.radius_search(src, tgt, 3.0, 1000)
    assert np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert a[2].max() > 20


def test_transform_matches_pcl_formula(oracle):
    src, _, T = synth.config1_plane_sphere(n_plane=400, n_sphere=100)
    assert np.array_equal(oracle.transform(src, T)[:, :3], synth.apply_T_like_pcl(src, T)[:, :3])


def test_voxel_grid(oracle):
    rng = np.random.default_rng(1)
    pts = np.ones((5000, 4), dtype=np.float32)
    pts[:, :3] = rng.uniform(-3, 3, (5000, 3))
    out, overflow = oracle.voxel_grid(pts, 0.5)
    assert not overflow and 0 < len(out) <= 12 ** 3
    # every centroid lies in a distinct voxel, ordered by ascending voxel id (x fastest)
    inv = np.float32(1.0) / np.float32(0.5)
    ijk = np.floor(out[:, :3] * inv).astype(np.int64)
    ijk -= np.floor(pts[:, :3].min(0) * inv).astype(np.int64)
    div = np.floor(pts[:, :3].max(0) * inv).astype(np.int64) - np.floor(pts[:, :3].min(0) * inv).astype(np.int64) + 1
    vid = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    assert np.all(np.diff(vid) > 0)
    # mass is conserved: the count-weighted mean of centroids is the cloud mean
    _, overflow = oracle.voxel_grid(pts * np.float32(1e4), 0.001)
    assert overflow


def test_outer_loop_semantics(oracle):
    src, tgt, _ = synth.config1_plane_sphere(seed=4, n_plane=500, n_sphere=300)
    opt = oracle.make_options(inner_kind=1)
    r0 = oracle.align(src, tgt, oracle.make_params(n_iter=0), opt)
    assert r0.n_total == 0                                            # n_iter == 0: align() is a no-op
    r3 = oracle.align(src, tgt, oracle.make_params(n_iter=3), opt)
    assert r3.n_total == 3
    full = oracle.align(src, tgt, oracle.make_params(), opt)
    drops = [s["cost_drop"] for s in full.stats]
    # terminates after the counter exceeded n_cost_drop_it: the last 6 recorded drops are all below threshold
    assert full.n_total < 1000 and all(d < 0.01 for d in drops[-6:])
    # history composes increments on the left: T_k = dT_k * T_{k-1}
    same = oracle.align(src, tgt, oracle.make_params(n_iter=full.n_total), opt)
    np.testing.assert_allclose(same.history, full.history, atol=0)
    # a perfectly aligned pair: zero cost -> NaN drop -> counter reset -> runs to n_iter
    z = oracle.align(tgt, tgt, oracle.make_params(n_iter=9, max_neighbours=1, radius=0.01), opt)
    assert z.n_total == 9 and np.isnan(z.stats[0]["cost_drop"])


def test_mse_helper(oracle):
    a = np.zeros((3, 4), dtype=np.float32)
    b = a.copy()
    b[:, 0] = [3, 0, 0]
    b[:, 1] = [4, 0, 0]
    assert oracle.calculate_mse(a, b) == pytest.approx(5.0 / 3.0)    # a mean of unsquared distances
