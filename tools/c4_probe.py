"""The 10M-point pair (BASELINE configs[3]) on ONE GPU under different environment switches: per-search launch times
(PPCR_TRACE_SEARCH), total, K.  The clouds are generated once and cached in /dev/shm.

    python tools/c4_probe.py "PPCR_Q_LEAVES=64" "PPCR_Q_LEAVES=4096" ...      (one subprocess per setting; "" = defaults)
"""
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/dev/shm/ppcr_c4.npy"

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    from probabilistic_point_clouds_registration_b200 import capi
    z = np.load(CACHE, mmap_mode="r")
    src, tgt = np.ascontiguousarray(z[0]), np.ascontiguousarray(z[1])
    params = capi.make_params(max_neighbours=10, radius=0.5, dof=5.0, n_iter=int(os.environ.get("C4_ITERS", "1000")))
    for rep in range(2):
        t0 = time.perf_counter()
        with capi.Registration(src, tgt, params, capi.make_options(driver=1, record_stage_times=True)) as reg:
            reg.align()
            stats = reg.iteration_stats()
            st = reg.stage_times()
        print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.1f} ms incl. H2D; {len(stats)} outer, "
              f"K={sum(s['n_correspondences'] for s in stats)}; search {st.search_ms:.1f} ms eval {st.eval_ms:.1f} ms", flush=True)
    sys.exit(0)

if not os.path.exists(CACHE):
    from probabilistic_point_clouds_registration_b200 import synth
    rings, az = int(os.environ.get("C4_RINGS", "320")), int(os.environ.get("C4_AZ", "31250"))
    src, tgt, _ = synth.lidar_pair(4, rings, az)
    np.save(CACHE, np.stack([src, tgt]))
for setting in sys.argv[1:] or [""]:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("=", 1)
        env[k] = v
    env["PPCR_TRACE_SEARCH"] = "1"
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
    times = [float(x) for x in re.findall(r"search launch\s+([0-9.]+) ms", r.stderr)]
    half = len(times) // 2
    print(f"== [{setting}] rc={r.returncode}")
    print(r.stdout.strip())
    if times:
        print("   searches of rep 1 (ms): " + " ".join(f"{t:.2f}" for t in times[half:]))
    if r.returncode:
        print(r.stderr[-2000:])
