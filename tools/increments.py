"""Per-outer-iteration increments of a workload (how far the cloud moves between searches)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from probabilistic_point_clouds_registration_b200 import capi
w = sys.argv[1] if len(sys.argv) > 1 else "c3"
src, tgt = bench.make_pair(w, 0)
with capi.Registration(src, tgt, capi.make_params(**bench.WORKLOADS[w]["params"])) as reg:
    reg.align()
    inc = reg.increment_history()
    st = reg.iteration_stats()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"increments_{w}.npy"), np.asarray(inc))
R = np.linalg.norm(src[:, :3], axis=1)
for k, T in enumerate(inc):
    ang = np.arccos(np.clip((np.trace(T[:3, :3]) - 1) / 2, -1, 1))
    p = src[::997, :3]
    mv = np.linalg.norm(p @ T[:3, :3].T + T[:3, 3] - p, axis=1)
    print(f"it {k:2d}: |dt|={np.linalg.norm(T[:3,3]):.2e} m  angle={ang:.2e} rad  point move median={np.median(mv):.2e} max={mv.max():.2e}  "
          f"lm={st[k]['lm_iterations']} drop={st[k]['cost_drop']:.4f}")
