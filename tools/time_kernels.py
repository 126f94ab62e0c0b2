"""Per-kernel timings of a bench workload on the state a registration ends in (ppcr_time_kernel).

    python tools/time_kernels.py [c3|c1|c5] [n_iter] [reps]
With PPCR_PROFILE_KERNEL=<which> and `ncu --profile-from-start off` only that kernel's launches are captured.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from probabilistic_point_clouds_registration_b200 import capi  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
src, tgt = bench.make_pair(workload, 0)
params = capi.make_params(n_iter=n_iter, **bench.WORKLOADS[workload]["params"])
names = {0: "search (cold bound)", 4: "search (moved + warm bound)", 1: "weights+normal eq+controller", 2: "cloud move",
         3: "target tree build", 5: "  weights+normal eq only (probe)", 6: "  ... + fold, no LM (probe)"}
leaf = int(os.environ.get("PPCR_LEAF", "0"))
with capi.Registration(src, tgt, params, capi.make_options(leaf_capacity=leaf)) as reg:
    reg.align()
    print(f"{workload}: {len(reg.iteration_stats())} outer iterations")
    for which in (0, 4, 1, 5, 6, 2, 3):
        ms, nbytes = reg.time_kernel(which, reps=reps, flush_l2=True)
        hot, _ = reg.time_kernel(which, reps=reps, flush_l2=False)
        print(f"  {names[which]:32s} {ms*1e3:9.1f} us  {nbytes/1e6:8.1f} MB algorithmic  {nbytes/ms/1e6:8.1f} GB/s   (L2 not flushed: {hot*1e3:.1f} us)")
