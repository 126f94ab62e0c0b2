"""One pair sharded over 2 GPUs (source slices + replicated target, moments exchanged peer-to-peer from inside the
evaluation kernel) against the single-GPU run.  Needs two B200s; skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_pair_on_two_gpus(capi):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GPU_OK" in r.stdout


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_sharded_pair_in_one_process(capi, n_ranks):
    """ppcr_align_sharded: the ranks are host threads of this process, the mailboxes are reached by plain peer access -- or, on
    a one-GPU box, all live on device 0 (the kernels of the ranks then exchange their moments through the same buffer, side by
    side on the same device): the whole in-kernel exchange runs wherever the suite runs.  Same iterations and pose as one handle."""
    import numpy as np
    import torch
    from helpers import pose_delta
    from probabilistic_point_clouds_registration_b200 import synth
    n_dev = torch.cuda.device_count()
    devices = [r % n_dev for r in range(n_ranks)]
    src, tgt, _ = synth.lidar_pair(51, 48, 900, yaw_deg=1.5, trans=(0.3, 0.05, 0.0))
    params = capi.make_params(max_neighbours=10, radius=0.8, dof=5.0)
    with capi.Registration(src, tgt, params) as reg:
        reg.align()
        ref = reg.transformation_history()
        k_ref = sum(s["n_correspondences"] for s in reg.iteration_stats())
    for exact in (False, True):
        if exact:
            with capi.Registration(src, tgt, params, capi.make_options(exact_weights=True)) as reg:
                reg.align()
                ref = reg.transformation_history()
                k_ref = sum(s["n_correspondences"] for s in reg.iteration_stats())
        hist, corr = capi.align_sharded(src, tgt, params, devices, capi.make_options(exact_weights=exact))
        assert len(hist) == len(ref)
        rot, tr = pose_delta(hist[-1], ref[-1])
        assert rot < 1e-7 and tr < 1e-7, (rot, tr)
        assert abs(corr - k_ref) <= 1e-5 * k_ref
    # a second registration re-uses the process-lifetime mailboxes (new stamp epoch): same bits as the first
    again, _ = capi.align_sharded(src, tgt, params, devices, capi.make_options(exact_weights=True))
    assert np.array_equal(again, hist)
    with pytest.raises(capi.PpcrError) as e:
        capi.align_sharded(src, tgt, capi.make_params(source_filter_size=0.1), devices)
    assert e.value.code == 4
