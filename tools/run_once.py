"""One registration of a bench workload through the C ABI (used under ncu; never a bench number).

    python tools/run_once.py [c3|c1|c5] [n_iter] [driver]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from probabilistic_point_clouds_registration_b200 import capi  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
driver = int(sys.argv[3]) if len(sys.argv) > 3 else 1
src, tgt = bench.make_pair(workload, 0)
params = capi.make_params(n_iter=n_iter, **bench.WORKLOADS[workload]["params"])
for rep in range(2):
    t0 = time.perf_counter()
    with capi.Registration(src, tgt, params, capi.make_options(driver=driver, record_stage_times=(driver == 1))) as reg:
        t1 = time.perf_counter()
        reg.align()
        t2 = time.perf_counter()
        stats = reg.iteration_stats()
        st = reg.stage_times()
    t3 = time.perf_counter()
    print(f"rep {rep}: ctor {1e3*(t1-t0):.2f} ms, align {1e3*(t2-t1):.2f} ms, teardown {1e3*(t3-t2):.2f} ms; "
          f"{len(stats)} outer, {sum(s['lm_iterations'] + 1 for s in stats)} evals, "
          f"K={sum(s['n_correspondences'] for s in stats)}; stage ms: search {st.search_ms:.2f} eval {st.eval_ms:.2f} "
          f"ctrl {st.controller_ms:.2f} transform {st.transform_ms:.2f}; launches {st.total_launches}")
