"""Weights + normal equations on the GPU vs the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from helpers import csr_from_rows, emu_normal_eq
from probabilistic_point_clouds_registration_b200 import synth

pytestmark = pytest.mark.gpu

IDENT = np.array([1.0, 0, 0, 0, 0, 0, 0])


def _golden_fixture():
    # test/ProbabilisticWeightsTest.cc:16-33: squared errors {1,1,1 | 1,4,9,16} on the pattern
    # row0 -> cols {0,2,3}, row1 -> cols {0,1,2,3}.  Realised geometrically: target j sits at distance
    # sqrt(e) from the row's source point, identity pose.
    src = np.array([[0, 0, 0, 1], [10, 0, 0, 1]], dtype=np.float32)
    tgt = np.array([[1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 1, 1],            # distance 1 from source 0
                    [11, 0, 0, 1], [12, 0, 0, 1], [13, 0, 0, 1], [14, 0, 0, 1]], dtype=np.float32)
    idx = np.array([[0, 1, 2, -1], [3, 4, 5, 6]], dtype=np.int32)
    cnt = np.array([3, 4], dtype=np.int32)
    return src, tgt, idx, cnt


@pytest.mark.parametrize("fast", [False, True])
def test_golden_t_distribution(capi, fast):
    src, tgt, idx, cnt = _golden_fixture()
    w, _ = capi.weights_normal_eq(src, tgt, idx, cnt, 5.0, IDENT, IDENT, dimension=1, fast_weights=fast)
    np.testing.assert_allclose(w[0, :3], [1 / 3] * 3, atol=1e-6)
    np.testing.assert_allclose(w[1], [0.7151351, 0.1412613, 0.0241258, 0.0047656], atol=1e-6)  # T_W:42-43


@pytest.mark.parametrize("fast", [False, True])
def test_golden_gaussian(capi, fast):
    src, tgt, idx, cnt = _golden_fixture()
    w, _ = capi.weights_normal_eq(src, tgt, idx, cnt, np.inf, IDENT, IDENT, dimension=1, fast_weights=fast)
    np.testing.assert_allclose(w[0, :3], [1 / 3] * 3, atol=1e-6)
    np.testing.assert_allclose(w[1], [0.805153702921689, 0.179654074677018, 0.0147469044726408,
                                      0.000445317928652638], atol=1e-6)  # T_W:59-60


@pytest.mark.parametrize("dof", [5.0, 1.5, np.inf])
@pytest.mark.parametrize("fast", [False, True])
def test_weights_match_oracle(capi, oracle, dof, fast):
    src, tgt, _ = synth.config1_plane_sphere(seed=5, n_plane=1500, n_sphere=1500)
    idx, _, cnt, _ = oracle.radius_search(src, tgt, 1.0, 20)
    pose_w = np.array([0.999, 0.02, -0.01, 0.03, 0.05, -0.02, 0.01])
    w, _ = capi.weights_normal_eq(src, tgt, idx, cnt, dof, pose_w, pose_w, fast_weights=fast)
    row_ptr, col = csr_from_rows(idx, cnt)
    # the oracle wants each row sorted by column; map its weights back to (row, column)
    _, ow = oracle.callback_weights(src, tgt, row_ptr, col, pose_w[:4], pose_w[4:], dof)
    worst = 0.0
    for i in range(len(src)):
        c = cnt[i]
        if c == 0:
            continue
        ref = dict(zip(col[row_ptr[i]:row_ptr[i + 1]].tolist(), ow[row_ptr[i]:row_ptr[i + 1]].tolist()))
        got = w[i, :c]
        exp = np.array([ref[j] for j in idx[i, :c]])
        worst = max(worst, np.max(np.abs(got - exp) / exp))
    assert worst < 1e-5, worst  # north_star: weights within 1e-5 relative
    if not fast:
        assert worst < 1e-11


@pytest.mark.parametrize("dof", [5.0, np.inf])
def test_normal_equations_match_host_logic(capi, emu, oracle, dof):
    """The eval kernel's moments, reduced and expanded like the controller does, equal the CPU build of the same
    headers to rounding, for different weight / residual poses (the one-iteration lag of the reference)."""
    src, tgt, _ = synth.config1_plane_sphere(seed=6, n_plane=2000, n_sphere=1000)
    idx, _, cnt, _ = oracle.radius_search(src, tgt, 1.0, 20)
    pose_w = np.array([1.0, 0.01, 0.02, -0.01, 0.01, 0.0, 0.02])
    pose_e = np.array([0.98, 0.03, -0.02, 0.05, 0.03, -0.04, 0.01])
    _, ne = capi.weights_normal_eq(src, tgt, idx, cnt, dof, pose_w, pose_e, want_weights=False)
    ref, _ = emu_normal_eq(emu, src, tgt, idx, cnt, dof, pose_w, pose_e)
    scale = np.abs(ref).max()
    assert np.max(np.abs(ne - ref)) / scale < 1e-12
    # the default float32 row arithmetic: same system to float32 rounding of the per-row sums
    _, nf = capi.weights_normal_eq(src, tgt, idx, cnt, dof, pose_w, pose_e, want_weights=False, fast_weights=True)
    assert np.max(np.abs(nf - ref)) / scale < 2e-6


@pytest.mark.parametrize("dof", [5.0, np.inf])
def test_normal_equations_match_oracle(capi, oracle, dof):
    """J^T W J, J^T W r and the cost against the ORACLE's own assembly (analytic Jacobian rows per correspondence, float64,
    oracle/ppcr_oracle.cpp NormalEqProblem) -- not against a CPU build of the product's headers.  The kernel reaches the same
    36 numbers through 24 source-point moments; they agree to rounding."""
    src, tgt, _ = synth.config1_plane_sphere(seed=6, n_plane=2000, n_sphere=1000)
    idx, _, cnt, _ = oracle.radius_search(src, tgt, 1.0, 20)
    row_ptr, col = csr_from_rows(idx, cnt)
    pose_w = np.array([1.0, 0.01, 0.02, -0.01, 0.01, 0.0, 0.02])
    pose_e = np.array([0.98, 0.03, -0.02, 0.05, 0.03, -0.04, 0.01])
    ref = oracle.normal_eq(src, tgt, row_ptr, col, dof, pose_w, pose_e)
    scale_h, scale_g = np.abs(ref[:28]).max(), np.abs(ref[28:35]).max()
    _, ne = capi.weights_normal_eq(src, tgt, idx, cnt, dof, pose_w, pose_e, want_weights=False)
    assert np.max(np.abs(ne[:28] - ref[:28])) / scale_h < 1e-11
    assert np.max(np.abs(ne[28:35] - ref[28:35])) / scale_g < 1e-10  # (the gradient is a sum of cancelling terms)
    assert abs(ne[35] - ref[35]) / ref[35] < 1e-12
    _, nf = capi.weights_normal_eq(src, tgt, idx, cnt, dof, pose_w, pose_e, want_weights=False, fast_weights=True)
    assert np.max(np.abs(nf[:28] - ref[:28])) / scale_h < 2e-6
    assert np.max(np.abs(nf[28:35] - ref[28:35])) / scale_g < 2e-5
    assert abs(nf[35] - ref[35]) / ref[35] < 2e-6


def test_rows_without_neighbours_contribute_nothing(capi):
    src, tgt, idx, cnt = _golden_fixture()
    cnt0 = np.array([0, 4], dtype=np.int32)
    _, ne_a = capi.weights_normal_eq(src, tgt, idx, cnt0, 5.0, IDENT, IDENT)
    _, ne_b = capi.weights_normal_eq(src[1:], tgt, idx[1:], cnt0[1:], 5.0, IDENT, IDENT)
    np.testing.assert_allclose(ne_a, ne_b, rtol=0, atol=0)
