"""align(): the whole outer loop on the GPU vs the oracle (src/prob_point_cloud_registration.cc:63-158)."""
import numpy as np
import pytest

from helpers import pose_delta
from probabilistic_point_clouds_registration_b200 import synth

pytestmark = pytest.mark.gpu

POSE_TOL_RAD = 1e-4   # north_star: final pose within 1e-4 rad and 1e-4 m of the reference's CPU result
POSE_TOL_M = 1e-4


def _run_both(capi, oracle, src, tgt, driver=0, inner_kind=1, exact=True, **kw):
    """exact=True selects the float64 weight arithmetic, which follows the oracle decision by decision;
    exact=False is the library default (float32 row arithmetic), held to the north_star tolerances."""
    gp = capi.make_params(**kw)
    op = oracle.make_params(**kw)
    with capi.Registration(src, tgt, gp, capi.make_options(driver=driver, exact_weights=exact)) as reg:
        reg.align()
        hist = reg.transformation_history()
        stats = reg.iteration_stats()
        moved = reg.filtered_source()
        done = reg.has_converged()
    ref = oracle.align(src, tgt, op, oracle.make_options(inner_kind=inner_kind), use_grid=True)
    return hist, stats, moved, done, ref


def _assert_parity(hist, stats, moved, ref):
    """Per outer iteration the association size, LM iteration counts and costs must agree exactly (costs to 1e-7)
    for as long as both sides take the same LM decisions.  Ceres' function-tolerance test is a threshold on a
    quantity that differs in the last bits between the two summation orders, so a long inner solve may stop one LM
    iteration apart; from there on the comparison is the north_star bar: same stopping iteration (+-1) and the
    final pose within 1e-4 rad / 1e-4 m."""
    lockstep = True
    for k in range(min(len(hist), ref.n_outer)):
        a, b = stats[k], ref.stats[k]
        if lockstep and (a["lm_iterations"], a["num_successful_steps"]) != (b["lm_iterations"], b["num_successful_steps"]):
            lockstep = False
            assert abs(a["lm_iterations"] - b["lm_iterations"]) <= 2, (k, a, b)
            assert a["n_correspondences"] == b["n_correspondences"], k
            np.testing.assert_allclose(a["initial_cost"], b["initial_cost"], rtol=1e-7)
            np.testing.assert_allclose(a["final_cost"], b["final_cost"], rtol=1e-4)
            continue
        if lockstep:
            assert a["n_correspondences"] == b["n_correspondences"], k
            np.testing.assert_allclose(a["initial_cost"], b["initial_cost"], rtol=1e-7)
            np.testing.assert_allclose(a["final_cost"], b["final_cost"], rtol=1e-7)
        else:
            assert abs(a["n_correspondences"] - b["n_correspondences"]) <= 1e-3 * b["n_correspondences"] + 2, k
            np.testing.assert_allclose(a["final_cost"], b["final_cost"], rtol=1e-3)
    if lockstep:
        assert len(hist) == ref.n_outer, (len(hist), ref.n_outer)
    else:
        assert abs(len(hist) - ref.n_outer) <= 1, (len(hist), ref.n_outer)
    rot, tr = pose_delta(hist[-1], ref.transformation)
    assert rot < POSE_TOL_RAD and tr < POSE_TOL_M, (rot, tr)
    # the moved float32 cloud: identical up to the last-bit effects of fp64 summation order
    tol = 1e-5 if lockstep else 2e-4 * max(1.0, float(np.max(np.abs(moved[:, :3]))))
    assert np.max(np.abs(moved[:, :3] - ref.filtered_source[:, :3])) < tol
    return lockstep


@pytest.mark.parametrize("driver", [1, 2])
@pytest.mark.parametrize("radius", [1.0, 3.0])
def test_config1_plane_sphere(capi, oracle, driver, radius):
    """BASELINE config 1: 10k-point plane+sphere, 10 deg / 0.1 m, t-distribution, struct and CLI radius."""
    src, tgt, _ = synth.config1_plane_sphere()
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, driver=driver, max_neighbours=20, dof=5.0,
                                              radius=radius)
    assert done
    _assert_parity(hist, stats, moved, ref)


def _assert_tolerance_parity(hist, stats, ref):
    """The default (float32 row arithmetic) path: the stopping iteration may move by one when a threshold test sits
    on the fence; the final pose must be within 1e-4 rad / 1e-4 m and the first association identical."""
    assert abs(len(hist) - ref.n_outer) <= 1, (len(hist), ref.n_outer)
    assert stats[0]["n_correspondences"] == ref.stats[0]["n_correspondences"]
    np.testing.assert_allclose(stats[0]["initial_cost"], ref.stats[0]["initial_cost"], rtol=1e-5)
    rot, tr = pose_delta(hist[-1], ref.transformation)
    assert rot < POSE_TOL_RAD and tr < POSE_TOL_M, (rot, tr)


@pytest.mark.parametrize("radius,dof", [(1.0, 5.0), (3.0, 5.0), (1.0, np.inf)])
def test_config1_default_weights_path(capi, oracle, radius, dof):
    """BASELINE config 1 through the library's default options."""
    src, tgt, _ = synth.config1_plane_sphere()
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, exact=False, max_neighbours=20, dof=dof,
                                              radius=radius)
    assert done
    _assert_tolerance_parity(hist, stats, ref)


@pytest.mark.parametrize("m,dof", [(5, 4.0), (33, 2.5), (64, 5.0), (128, np.inf), (7, 6.0)])
def test_default_path_weight_models_and_row_lengths(capi, oracle, m, dof):
    """Every instantiation of the default evaluation's row loop -- t with a half-integer exponent (dof 4), with a real
    one (dof 2.5), with exponent 4 (dof 5), Gaussian, integer exponent (dof 6: (6+3)/2 is half-integer, dof 7 would be 5)
    -- at row lengths that are not multiples of the gather batch, exceed one warp of bulk-copy lanes (m > 32) and
    reach the maximum (128)."""
    src, tgt, _ = synth.config1_plane_sphere(seed=31, n_plane=900, n_sphere=700)
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, exact=False, max_neighbours=m, dof=dof,
                                              radius=1.0)
    assert done
    _assert_tolerance_parity(hist, stats, ref)


def test_config2_gaussian_outliers_default_path(capi, oracle):
    """BASELINE config 2, reduced, through the library defaults: 20% outliers, Gaussian weights, both voxel filters."""
    src, tgt, _ = synth.lidar_pair(2, 32, 700, outlier_frac=0.2)
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, exact=False, max_neighbours=20, dof=np.inf,
                                              radius=3.0, source_filter_size=0.25, target_filter_size=0.25)
    assert done
    assert len(moved) == ref.n_filtered_src < len(src)
    _assert_tolerance_parity(hist, stats, ref)


def test_config2_full_size_pose_parity(capi, oracle):
    """BASELINE config 2 at its stated size: ~100k rays (64 rings x 1563 azimuths), 20% uniform outliers, the flags
    `-u -s 0.05 -t 0.05` with everything else at the CLI defaults (m 20, r 3).  The scene is bounded (60 x 60 x 20 m box)
    so that the 0.05 m voxel grid does not overflow int32 and both filters really run."""
    src, tgt, _ = synth.config2_lidar_outliers()
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, exact=False, max_neighbours=20, dof=np.inf,
                                              radius=3.0, source_filter_size=0.05, target_filter_size=0.05)
    assert done
    assert len(moved) == ref.n_filtered_src < len(src)
    _assert_tolerance_parity(hist, stats, ref)


def test_config5_pairs_through_the_batch_entry(capi, oracle):
    """BASELINE config 5 at its stated size (120k-point pairs, seeds 1000.., CLI defaults m 20 / r 3 / dof 5) through
    ppcr_align_batch with several lanes: every pair against the oracle run on its own."""
    idx = [0, 1, 2, 3]
    pairs = [synth.config5_pair(i)[:2] for i in idx]
    kw = dict(max_neighbours=20, dof=5.0, radius=3.0)
    T, n_outer, corr = capi.align_batch(pairs, capi.make_params(**kw), slots=3)
    for k, (s, t) in enumerate(pairs):
        ref = oracle.align(s, t, oracle.make_params(**kw), oracle.make_options(inner_kind=1), use_grid=True)
        assert abs(int(n_outer[k]) - ref.n_outer) <= 1, (k, n_outer[k], ref.n_outer)
        if int(n_outer[k]) == ref.n_outer:
            want = sum(x["n_correspondences"] for x in ref.stats)
            assert abs(int(corr[k]) - want) <= 1e-4 * want, (k, corr[k], want)
        rot, tr = pose_delta(T[k], ref.transformation)
        assert rot < POSE_TOL_RAD and tr < POSE_TOL_M, (k, rot, tr)


def test_config3_full_size_pose_parity(capi, oracle):
    """BASELINE config 3 at full size (1M-point pair, -m 10 -r 0.5 -d 5), library defaults, against the oracle run
    to its own stopping rule: final pose within 1e-4 rad / 1e-4 m."""
    src, tgt, _ = synth.config3_lidar_1m()
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, exact=False, max_neighbours=10, dof=5.0,
                                              radius=0.5)
    assert done
    _assert_tolerance_parity(hist, stats, ref)


def test_gaussian_with_voxel_filters(capi, oracle):
    """BASELINE config 2, reduced: LiDAR-like pair, 20% outliers, -u, voxel filter on both clouds."""
    src, tgt, _ = synth.lidar_pair(2, 32, 700, outlier_frac=0.2)
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, max_neighbours=20, dof=np.inf, radius=3.0,
                                              source_filter_size=0.25, target_filter_size=0.25)
    assert len(moved) == ref.n_filtered_src < len(src)
    _assert_parity(hist, stats, moved, ref)


def test_faithful_oracle_small(capi, oracle):
    """Against the dual-number + dense-QR oracle (the closest restatement of Ceres AutoDiff + DENSE_QR)."""
    src, tgt, _ = synth.config1_plane_sphere(seed=11, n_plane=700, n_sphere=500)
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, inner_kind=0, max_neighbours=20, dof=5.0,
                                              radius=1.0)
    _assert_parity(hist, stats, moved, ref)


def test_has_converged_semantics(capi):
    """hasConverged() mutates (registration.cc:138-158): n_iter==0 is converged at once; with the counter at
    zero it takes n_cost_drop_it + 2 calls of an idle handle to report convergence."""
    src, tgt, _ = synth.config1_plane_sphere(n_plane=300, n_sphere=300)
    with capi.Registration(src, tgt, capi.make_params(n_iter=0)) as reg:
        assert reg.has_converged()
        reg.align()
        assert len(reg.transformation_history()) == 0
        with pytest.raises(IndexError):
            reg.transformation()
    with capi.Registration(src, tgt, capi.make_params(n_iter=50, n_cost_drop_it=2)) as reg:
        assert [reg.has_converged() for _ in range(4)] == [False, False, False, True]


def test_n_iter_cap_and_history(capi, oracle):
    src, tgt, _ = synth.config1_plane_sphere(n_plane=600, n_sphere=400)
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, max_neighbours=8, dof=5.0, radius=0.7, n_iter=3)
    assert len(hist) == 3 and done
    _assert_parity(hist, stats, moved, ref)


def test_zero_cost_runs_all_iterations(capi, oracle):
    """A perfectly aligned noiseless pair has initial cost 0 -> NaN cost drop -> the counter resets and the loop
    only stops at n_iter."""
    tgt = synth.reference_test_cloud()
    hist, stats, moved, done, ref = _run_both(capi, oracle, tgt, tgt, max_neighbours=1, dof=5.0, radius=0.2, n_iter=12)
    assert len(hist) == ref.n_outer == 12
    np.testing.assert_allclose(hist[-1], np.eye(4), atol=1e-12)


def test_batch_matches_single(capi):
    pairs = [synth.lidar_pair(100 + i, 16, 400, random_motion=(2.0, 0.5))[:2] for i in range(5)]
    params = capi.make_params(max_neighbours=12, radius=2.0, n_iter=30)
    T, n_outer, corr = capi.align_batch(pairs, params, slots=3)
    for i, (s, t) in enumerate(pairs):
        with capi.Registration(s, t, params) as reg:
            reg.align()
            h = reg.transformation_history()
            st = reg.iteration_stats()
        assert n_outer[i] == len(h)
        assert corr[i] == sum(x["n_correspondences"] for x in st)
        np.testing.assert_allclose(T[i], h[-1], rtol=0, atol=1e-12)


def test_batch_over_a_device_list(capi):
    """ppcr_align_batch_devices (device_ids[], n_dev): lanes on every listed device draw from one counter; the
    result of a pair does not depend on the device or lane that ran it.  On a one-GPU box the list names device 0 twice."""
    import torch
    n_dev = torch.cuda.device_count()
    devices = list(range(min(n_dev, 4))) if n_dev > 1 else [0, 0]
    pairs = [synth.lidar_pair(200 + i, 16, 400, random_motion=(2.0, 0.5))[:2] for i in range(7)]
    params = capi.make_params(max_neighbours=12, radius=2.0, n_iter=30)
    T1, n1, c1 = capi.align_batch(pairs, params, slots=2)
    T2, n2, c2 = capi.align_batch(pairs, params, slots=2, devices=devices)
    assert np.array_equal(n1, n2) and np.array_equal(c1, c2)
    np.testing.assert_array_equal(T1, T2)
    dev_opt = capi.make_options(input_on_device=True)
    with pytest.raises(capi.PpcrError) as e:
        capi.align_batch([(0, 0, 0, 0)], params, dev_opt, devices=[0, 0])
    assert e.value.code == 1


@pytest.mark.parametrize("max_neighbours", [0, -3, 500])
def test_unlimited_neighbour_sets(capi, oracle, max_neighbours):
    """max_neighbours <= 0 (pcl: every target within the radius; reachable with `-m 0`, registration.cc:74-75) and values above
    the row capacity of 128: rows hold ALL in-radius targets as long as no row reaches 128 of them."""
    src, tgt, _ = synth.config1_plane_sphere(seed=5, n_plane=1500, n_sphere=1000)
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, tgt, max_neighbours=max_neighbours, dof=5.0, radius=0.7)
    assert done and 20 < stats[0]["n_correspondences"] / len(src) < 128  # well beyond the default of 20 per row
    _assert_parity(hist, stats, moved, ref)


def test_unlimited_neighbour_sets_small_target_and_overflow(capi, oracle):
    src, tgt, _ = synth.config1_plane_sphere(seed=6, n_plane=400, n_sphere=300)
    # a target of fewer points than the row capacity: "all of them" fits whatever the radius
    small = tgt[::7]
    assert len(small) <= 128
    hist, stats, moved, done, ref = _run_both(capi, oracle, src, small, max_neighbours=0, dof=5.0, radius=50.0, n_iter=4)
    assert stats[0]["n_correspondences"] == len(src) * len(small)
    _assert_parity(hist, stats, moved, ref)
    # a row that would need more than 128 neighbours stops the registration instead of being truncated silently
    with capi.Registration(src, tgt, capi.make_params(max_neighbours=0, radius=3.0)) as reg:
        with pytest.raises(capi.PpcrError) as e:
            reg.align()
        assert e.value.code == 4 and "128" in str(e.value)


def test_histories_are_bit_identical_from_run_to_run(capi, monkeypatch):
    """Every search path leaves a row in a reproducible order and every sum has a fixed order: the pose history of a pair is
    the same bits on every run -- also with a candidate list so short, or so few leaves per query, that most queries take the
    one-by-one fallback (whose pruning bound depends on which candidates arrived first) -- alone or in a batch of lanes."""
    src, tgt, _ = synth.lidar_pair(41, 48, 900, yaw_deg=1.5, trans=(0.3, 0.05, 0.0))
    params = capi.make_params(max_neighbours=10, radius=0.8, dof=5.0)

    def run():
        with capi.Registration(src, tgt, params) as reg:
            reg.align()
            return reg.transformation_history()

    base = run()
    assert len(base) > 3
    for _ in range(2):
        assert np.array_equal(run(), base)
    # the queue path and its fallbacks (no chunk takes the heap walk: its rows are in heap order, equally reproducible)
    monkeypatch.setenv("PPCR_Q_HEAVY", "1e9")
    queued = run()
    for env in ({"PPCR_Q_CAND": "12"}, {"PPCR_Q_LEAVES": "3"}):
        with monkeypatch.context() as mp:
            for k, v in env.items():
                mp.setenv(k, v)
            assert np.array_equal(run(), queued), env
    monkeypatch.delenv("PPCR_Q_HEAVY")
    T, n_outer, _ = capi.align_batch([(src, tgt)] * 7, params, slots=4)
    assert np.all(n_outer == len(base))
    for k in range(7):
        assert np.array_equal(T[k], base[-1])
