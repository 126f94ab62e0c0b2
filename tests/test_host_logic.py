"""The product's scalar logic (csrc/ppcr_lm.h + csrc/ppcr_eval.h, compiled for the CPU by tests/emu) against
the oracle.  These are the same sources k_controller / k_eval compile for sm_100a."""
import numpy as np
import pytest

from helpers import emu_align, pose_delta
from probabilistic_point_clouds_registration_b200 import synth


@pytest.mark.parametrize("dof,radius,kind", [(5.0, 1.0, 0), (np.inf, 1.0, 0), (5.0, 3.0, 1)])
def test_controller_follows_the_restated_ceres_path(emu, oracle, dof, radius, kind):
    src, tgt, _ = synth.config1_plane_sphere(n_plane=700, n_sphere=500)
    oracle.nonmonotonic_steps(reset=True)
    ref = oracle.align(src, tgt, oracle.make_params(max_neighbours=20, dof=dof, radius=radius),
                       oracle.make_options(inner_kind=kind), use_grid=False)
    # Every one of these runs ACCEPTS non-monotonic steps (the weights are refreshed between the evaluation of x and of the
    # candidate, so the cost often rises on an accepted step).  After such a step Ceres' update_state_every_iteration hands the
    # callback its lowest-cost iterate, not x (oracle/ppcr_oracle.cpp minimise(), csrc/ppcr_lm.h step_prepare): the lock-step
    # comparison below only holds if both sides refresh the weights there.
    assert oracle.nonmonotonic_steps() > 0
    n, hist, stats, moved = emu_align(emu, src, tgt, 20, dof, radius)
    assert n == ref.n_outer
    for a, b in zip(stats, ref.stats):
        assert a["n_correspondences"] == b["n_correspondences"]
        assert a["lm_iterations"] == b["lm_iterations"]
        assert a["num_successful_steps"] == b["num_successful_steps"]
        np.testing.assert_allclose([a["initial_cost"], a["final_cost"]], [b["initial_cost"], b["final_cost"]], rtol=1e-8)
    np.testing.assert_allclose(hist, ref.history, atol=1e-8)
    assert np.max(np.abs(moved - ref.filtered_source)) < 1e-6


def test_fast_weights_stay_within_tolerance(emu, oracle):
    src, tgt, _ = synth.config1_plane_sphere(n_plane=600, n_sphere=400)
    ref = oracle.align(src, tgt, oracle.make_params(), oracle.make_options(inner_kind=1), use_grid=False)
    n, hist, _, _ = emu_align(emu, src, tgt, 20, 5.0, 1.0, fast=1)
    assert abs(n - ref.n_outer) <= 1
    rot, tr = pose_delta(hist[-1], ref.history[min(n, ref.n_outer) - 1])
    assert rot < 1e-4 and tr < 1e-4


def test_n_iter_zero_and_counter(emu):
    src, tgt, _ = synth.config1_plane_sphere(n_plane=100, n_sphere=100)
    n, hist, _, moved = emu_align(emu, src, tgt, 20, 5.0, 1.0, n_iter=0)
    assert n == 0 and np.array_equal(moved, src)
