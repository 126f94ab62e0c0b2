"""Committed fixtures (tests/golden, written by tests/golden/make_golden.py) against the oracle on CPU and against the
CUDA path on the GPU."""
import json
import os

import numpy as np
import pytest

from helpers import pose_delta
from probabilistic_point_clouds_registration_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = (("t5_r1", 5.0, 1.0), ("gauss_r1", np.inf, 1.0), ("t5_r3", 5.0, 3.0))


def _weights():
    return json.load(open(os.path.join(GOLDEN, "reference_weights.json")))


def test_reference_weight_vectors_oracle(oracle):
    g = _weights()
    row_ptr = np.cumsum([0] + [len(r) for r in g["pattern_rows"]])
    for model in ("t_distribution", "gaussian"):
        dof = np.inf if g[model]["dof"] == "inf" else float(g[model]["dof"])
        w = oracle.update_weights(row_ptr, np.array(g["squared_errors"], dtype=float), dof, g["dimension"])
        np.testing.assert_allclose(w, g[model]["row0"] + g[model]["row1"], atol=g[model]["tol"])


def test_reference_fixture_matches_generator():
    f = np.load(os.path.join(GOLDEN, "reference_fixture.npz"))
    assert np.array_equal(f["source"], synth.reference_test_cloud())
    np.testing.assert_allclose(f["T"][:3, 3], [2.35688666, 0.83371773, 0.0], atol=1e-8)


@pytest.mark.parametrize("name,dof,radius", CASES)
def test_oracle_reproduces_its_committed_outputs(oracle, name, dof, radius):
    f = np.load(os.path.join(GOLDEN, "oracle_c1_small.npz"))
    s, t, _ = synth.config1_plane_sphere(seed=int(f["seed"]), n_plane=int(f["n_plane"]), n_sphere=int(f["n_sphere"]))
    r = oracle.align(s, t, oracle.make_params(max_neighbours=20, dof=dof, radius=radius), oracle.make_options(inner_kind=1))
    assert np.array_equal(np.array([x["n_correspondences"] for x in r.stats]), f[name + "_K"])
    np.testing.assert_allclose(r.history, f[name + "_history"], atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name,dof,radius", CASES)
def test_cuda_path_against_committed_oracle_outputs(capi, name, dof, radius):
    f = np.load(os.path.join(GOLDEN, "oracle_c1_small.npz"))
    s, t, _ = synth.config1_plane_sphere(seed=int(f["seed"]), n_plane=int(f["n_plane"]), n_sphere=int(f["n_sphere"]))
    for exact in (True, False):
        with capi.Registration(s, t, capi.make_params(max_neighbours=20, dof=dof, radius=radius),
                               capi.make_options(exact_weights=exact)) as reg:
            reg.align()
            hist = reg.transformation_history()
            K = np.array([x["n_correspondences"] for x in reg.iteration_stats()])
        want = f[name + "_history"]
        assert abs(len(hist) - len(want)) <= (0 if exact else 1)
        assert K[0] == f[name + "_K"][0]
        if exact:
            assert np.array_equal(K, f[name + "_K"])
        rot, tr = pose_delta(hist[-1], want[-1])
        assert rot < 1e-4 and tr < 1e-4


@pytest.mark.gpu
def test_reference_weight_vectors_cuda(capi):
    g = _weights()
    # squared errors realised geometrically: target j at distance sqrt(e) from the row's source point
    src = np.array([[0, 0, 0, 1], [10, 0, 0, 1]], dtype=np.float32)
    tgt = np.array([[1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 1, 1], [11, 0, 0, 1], [12, 0, 0, 1], [13, 0, 0, 1], [14, 0, 0, 1]],
                   dtype=np.float32)
    idx = np.array([[0, 1, 2, -1], [3, 4, 5, 6]], dtype=np.int32)
    cnt = np.array([3, 4], dtype=np.int32)
    ident = np.array([1.0, 0, 0, 0, 0, 0, 0])
    for model in ("t_distribution", "gaussian"):
        dof = np.inf if g[model]["dof"] == "inf" else float(g[model]["dof"])
        for fast in (False, True):
            w, _ = capi.weights_normal_eq(src, tgt, idx, cnt, dof, ident, ident, dimension=g["dimension"], fast_weights=fast)
            np.testing.assert_allclose(w[0, :3], g[model]["row0"], atol=g[model]["tol"])
            np.testing.assert_allclose(w[1], g[model]["row1"], atol=g[model]["tol"])
