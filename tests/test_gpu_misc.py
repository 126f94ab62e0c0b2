"""Voxel filter, rigid transform and API error behaviour on the GPU."""
import numpy as np
import pytest

from probabilistic_point_clouds_registration_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("leaf", [0.05, 0.3, 2.0])
def test_voxel_filter_matches_oracle(capi, oracle, leaf):
    src, _, _ = synth.lidar_pair(4, 32, 800, outlier_frac=0.1)
    got = capi.voxel_filter(src, leaf)
    ref, overflow = oracle.voxel_grid(src, leaf)
    assert not overflow
    assert got.shape == ref.shape
    # same voxel set, same order; centroids accumulate in float32 in point-index order on both sides
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_voxel_filter_overflow_returns_input(capi, oracle):
    src, _, _ = synth.lidar_pair(4, 8, 100)
    big = src.copy()
    big[:, :3] *= 1000.0
    got = capi.voxel_filter(big, 0.01)   # (range / leaf)^3 overflows int32: PCL warns and returns the input
    ref, overflow = oracle.voxel_grid(big, 0.01)
    assert overflow and np.array_equal(got, big)


def test_transform_bit_exact(capi, oracle):
    src, _, T = synth.config1_plane_sphere(n_plane=5000, n_sphere=5000)
    got = capi.transform(src, T)
    ref = oracle.transform(src, T)
    assert np.array_equal(got[:, :3].view(np.uint32), ref[:, :3].view(np.uint32))


def test_invalid_arguments(capi):
    c = np.ones((8, 4), dtype=np.float32)
    with pytest.raises(capi.PpcrError) as e:
        capi.Registration(c, c, capi.make_params(radius=-1.0))
    assert e.value.code == 1 and "radius" in str(e.value)
    with pytest.raises(capi.PpcrError):
        capi.Registration(c, c, capi.make_params(max_neighbours=0))
    bad = c.copy()
    bad[3, 1] = np.nan
    with pytest.raises(capi.PpcrError):
        capi.Registration(c, bad, capi.make_params())


def test_target_is_filtered_and_source_copy_is_moved(capi):
    src, tgt, _ = synth.lidar_pair(5, 16, 300)
    with capi.Registration(src, tgt, capi.make_params(target_filter_size=0.5, n_iter=2)) as reg:
        ft = reg.filtered_target()
        assert 0 < len(ft) < len(tgt)           # what the reference writes back into the caller's target cloud
        before = reg.filtered_source().copy()
        reg.align()
        after = reg.filtered_source()
    assert len(before) == len(src) and not np.array_equal(before, after)
