"""The closest-point metric helpers (utilities.hpp:28-234) on the GPU vs the oracle's restatement."""
import numpy as np
import pytest

from probabilistic_point_clouds_registration_b200 import synth

pytestmark = pytest.mark.gpu

SUMS = ("average_closest_distance", "sum_squared_error", "robust_sum_squared_error", "robust_sum_squared_error_factor",
        "robust_averaged_sum_squared_error")
EXACT = ("median_closest_distance", "robust_median_closest_distance", "n_filtered", "n_filtered_factor")


def _compare(got, ref):
    for k in SUMS:  # double sums of the same float32 distances, added in another order: rounding only
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-12, err_msg=k)
    for k in EXACT:  # picked elements / counts: exact
        assert got[k] == ref[k] or (np.isnan(got[k]) and np.isnan(ref[k])), (k, got[k], ref[k])


@pytest.mark.parametrize("n1,n2,factor", [(3000, 2500, 3.0), (2999, 4000, 2.0), (64, 5000, 1.5)])
def test_metrics_match_oracle(capi, oracle, n1, n2, factor):
    src, tgt, _ = synth.config1_plane_sphere(seed=11, n_plane=2500, n_sphere=2500)
    a, b = src[:n1], tgt[:n2]
    got, d2 = capi.closest_point_metrics(a, b, factor)
    ref, rd2 = oracle.closest_metrics(a, b, factor)
    assert np.array_equal(d2.view(np.uint32), rd2.view(np.uint32))  # nearestKSearch(k = 1) squared distances, bit-exact
    _compare(got, ref)


def test_lidar_scan_against_itself_moved(capi, oracle):
    src, tgt, _ = synth.lidar_pair(7, 32, 600, outlier_frac=0.1)
    got, d2 = capi.closest_point_metrics(src, tgt)
    ref, rd2 = oracle.closest_metrics(src, tgt)
    assert np.array_equal(d2.view(np.uint32), rd2.view(np.uint32))
    _compare(got, ref)
    # a cloud against itself: every distance is zero, the windows hold everything, the sums vanish
    same, dz = capi.closest_point_metrics(tgt, tgt)
    assert not dz.any() and same["sum_squared_error"] == 0.0 and same["n_filtered"] == len(tgt)


def test_small_clouds_follow_the_reference_index_rules(capi, oracle):
    rng = np.random.default_rng(3)
    b = np.zeros((50, 4), dtype=np.float32)
    b[:, :3] = rng.normal(size=(50, 3))
    for n1 in (1, 2, 3, 4, 9, 10, 11, 25):
        a = np.zeros((n1, 4), dtype=np.float32)
        a[:, :3] = rng.normal(size=(n1, 3))
        got, _ = capi.closest_point_metrics(a, b)
        ref, _ = oracle.closest_metrics(a, b)
        _compare(got, ref)
        if n1 < 10:  # fewer than 10 inliers: the robust sums return DBL_MAX (utilities.hpp:97-99)
            assert got["robust_sum_squared_error"] == np.finfo(np.float64).max
    with pytest.raises(capi.PpcrError):
        capi.closest_point_metrics(b[:0], b)
    bad = b.copy()
    bad[7, 2] = np.inf
    with pytest.raises(capi.PpcrError):
        capi.closest_point_metrics(bad, b)
