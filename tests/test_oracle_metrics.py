"""CPU checks of oracle functions that have no reference-held vector: the closest-point metrics against an independent
numpy computation, and the normal equations against a finite-difference gradient."""
import numpy as np

from helpers import csr_from_rows
from probabilistic_point_clouds_registration_b200 import synth


def test_closest_metrics_against_numpy(oracle):
    rng = np.random.default_rng(5)
    a = np.zeros((401, 4), dtype=np.float32)
    b = np.zeros((300, 4), dtype=np.float32)
    a[:, :3] = rng.normal(size=(401, 3))
    b[:, :3] = rng.normal(size=(300, 3))
    m, d2 = oracle.closest_metrics(a, b, 2.5)
    d = a[:, None, :3] - b[None, :, :3]
    dd = ((d[..., 0] * d[..., 0]) + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]  # float32, FLANN's order
    nn = dd.min(axis=1)
    assert np.array_equal(nn, d2)
    s = np.sort(nn)
    s64 = s.astype(np.float64)
    assert m["sum_squared_error"] == float(np.cumsum(nn.astype(np.float64))[-1])
    assert m["average_closest_distance"] == m["sum_squared_error"] / len(a)
    med = float(s64[(len(s) + 1) // 2])  # odd size: the reference takes element (n + 1) / 2
    inl = s64[(s64 <= med * 3) & (s64 >= med / 3)]
    np.testing.assert_allclose(m["robust_sum_squared_error"], inl.sum(), rtol=1e-13)
    assert m["n_filtered"] == len(inl)
    np.testing.assert_allclose(m["robust_averaged_sum_squared_error"], inl.sum() / len(inl), rtol=1e-13)
    inl2 = s64[(s64 <= med * 2.5) & (s64 >= med / 2.5)]
    np.testing.assert_allclose(m["robust_sum_squared_error_factor"], inl2.sum(), rtol=1e-13)
    assert m["median_closest_distance"] == med
    filt = s[(s64 <= med * 3) & (s64 >= med / 3.0)]
    k = len(filt)
    fm = float(filt[(k + 1) // 2]) if k % 2 else float(np.float32(filt[k // 2] + filt[k // 2 + 1])) / 2.0
    assert m["robust_median_closest_distance"] == fm / k
    # even size: mean of elements n/2 and n/2 + 1
    m2, d2b = oracle.closest_metrics(a[:400], b)
    s2 = np.sort(d2b)
    assert m2["median_closest_distance"] == float(np.float32(s2[200] + s2[201])) / 2.0


def test_normal_equations_against_finite_differences(oracle):
    src, tgt, _ = synth.config1_plane_sphere(seed=8, n_plane=300, n_sphere=200)
    idx, _, cnt, _ = oracle.radius_search(src, tgt, 1.0, 8)
    row_ptr, col = csr_from_rows(idx, cnt)
    xw = np.array([1.0, 0.01, 0.02, -0.01, 0.01, 0.0, 0.02])
    xe = np.array([0.98, 0.03, -0.02, 0.05, 0.03, -0.04, 0.01])
    ne = oracle.normal_eq(src, tgt, row_ptr, col, 5.0, xw, xe)
    H = np.zeros((7, 7))
    H[np.triu_indices(7)] = ne[:28]
    H = H + np.triu(H, 1).T
    g, cost = ne[28:35], ne[35]
    _, w = oracle.callback_weights(src, tgt, row_ptr, col, xw[:4], xw[4:], 5.0)

    def cost_at(x):  # 1/2 sum w |y - (R(q/|q|) x + t)|^2 with the weights held fixed
        q = x[:4] / np.linalg.norm(x[:4])
        a, b = q[0], q[1:]
        tot = 0.0
        for i in range(len(src)):
            p = src[i, :3].astype(np.float64)
            rp = p + 2 * a * np.cross(b, p) + 2 * np.cross(b, np.cross(b, p)) + x[4:]
            for k in range(row_ptr[i], row_ptr[i + 1]):
                r = tgt[col[k], :3].astype(np.float64) - rp
                tot += 0.5 * w[k] * (r @ r)
        return tot

    np.testing.assert_allclose(cost, cost_at(xe), rtol=1e-12)
    num = np.zeros(7)
    for p in range(7):
        h = 1e-6
        e = np.zeros(7)
        e[p] = h
        num[p] = (cost_at(xe + e) - cost_at(xe - e)) / (2 * h)
    np.testing.assert_allclose(g, num, rtol=2e-6, atol=1e-7 * np.abs(num).max())
    assert np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H) > -1e-9 * np.abs(H).max())
    np.testing.assert_allclose(np.diag(H)[4:], w.sum(), rtol=1e-12)  # H_tt = (sum of the weights) I
