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
    # ppcr_transform_ex: on a given device and stream, in place on a device-resident cloud
    import torch
    stream = torch.cuda.Stream()
    d = torch.from_numpy(src).cuda()
    stream.wait_stream(torch.cuda.current_stream())
    capi.transform(d.data_ptr(), T, capi.make_options(input_on_device=True, stream=stream.cuda_stream), n_points=len(src))
    assert np.array_equal(d.cpu().numpy()[:, :3].view(np.uint32), ref[:, :3].view(np.uint32))
    got = capi.transform(src, T, capi.make_options(device=torch.cuda.device_count() - 1))
    assert np.array_equal(got[:, :3].view(np.uint32), ref[:, :3].view(np.uint32))


def test_invalid_arguments(capi):
    c = np.ones((8, 4), dtype=np.float32)
    with pytest.raises(capi.PpcrError) as e:
        capi.Registration(c, c, capi.make_params(radius=-1.0))
    assert e.value.code == 1 and "radius" in str(e.value)
    with pytest.raises(capi.PpcrError):
        capi.Registration(c, c, capi.make_params(dof=0.0))
    with capi.Registration(c, c, capi.make_params(max_neighbours=0)):  # pcl's "all in-radius targets": accepted (rows of 128)
        pass
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


def test_replay_metrics_match_a_host_replay(capi):
    """ppcr_replay_metrics = the reference's per-iteration bookkeeping (registration.cc:110-122 with calculateMSE,
    utilities.hpp:16-26) on the device: the cloud moved by every increment like pcl::transformPointCloud (bit-exact), and
    the two mean distances after each move (float32 distances, summed in double)."""
    src, tgt, T = synth.config1_plane_sphere(seed=5, n_plane=1500, n_sphere=1500)
    gt = synth.apply_T_like_pcl(src, T)
    with capi.Registration(src, tgt, capi.make_params(max_neighbours=10, radius=1.0)) as reg:
        reg.align()
        inc = reg.increment_history()
        moved, mse_gt, mse_prev = reg.replay_metrics(src, gt)
        half = len(inc) // 2
        part1, g1, p1 = reg.replay_metrics(src, gt, first=0, count=half)
        part2, g2, p2 = reg.replay_metrics(part1, None, first=half)
    assert len(inc) >= 3 and len(mse_gt) == len(inc)
    cur = src.copy()
    for k, Tk in enumerate(inc):
        new = synth.apply_T_like_pcl(cur, Tk)

        def mean_dist(a, b):
            d = a[:, :3] - b[:, :3]
            return float(np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float64).mean())

        np.testing.assert_allclose(mse_gt[k], mean_dist(new, gt), rtol=1e-12)
        np.testing.assert_allclose(mse_prev[k], mean_dist(new, cur), rtol=1e-12)
        cur = new
    assert np.array_equal(moved[:, :3], cur[:, :3])
    # a replay in two parts continues where the first stopped; without a ground truth that column is zero
    assert np.array_equal(part2[:, :3], moved[:, :3])
    assert np.array_equal(np.concatenate([g1, g2]), np.concatenate([mse_gt[:half], np.zeros(len(inc) - half)]))
    assert np.array_equal(np.concatenate([p1, p2]), mse_prev)
    assert mse_gt[-1] < mse_gt[0]


def test_page_locked_host_memory(capi):
    """ppcr_host_alloc / ppcr_host_free (the storage of the PCL stand-in's clouds): blocks it handed out are recognised and freed,
    foreign pointers are refused, and a cloud living in such a block registers like any other."""
    import ctypes as C
    L = capi.lib()
    n = 5000
    p = L.ppcr_host_alloc(n * 16)
    assert p
    buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, 4))
    src, tgt, _ = synth.config1_plane_sphere(seed=9, n_plane=n // 2, n_sphere=n - n // 2)
    buf[:] = tgt
    with capi.Registration(src, buf, capi.make_params(max_neighbours=8, radius=0.6)) as a, \
            capi.Registration(src, tgt, capi.make_params(max_neighbours=8, radius=0.6)) as b:
        a.align()
        b.align()
        assert np.array_equal(a.transformation_history(), b.transformation_history())
    del buf
    assert L.ppcr_host_free(p) == 1
    assert L.ppcr_host_free(p) == 0                       # already gone
    other = np.zeros(4, dtype=np.float32)
    assert L.ppcr_host_free(other.ctypes.data) == 0       # not ours: the caller frees it its own way
    assert not L.ppcr_host_alloc(0)
