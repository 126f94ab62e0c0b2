"""Synthetic clouds for the BASELINE.json configs and the reference's own test fixture.

All generators are numpy.random.Generator(PCG64(seed)) driven and return float32 [N,4] clouds
(x, y, z, 1) -- the 16-byte pcl::PointXYZ record -- so the same bits feed the CPU checker and the GPU.
"""
from __future__ import annotations

import numpy as np


def _xyzw(p):
    out = np.ones((len(p), 4), dtype=np.float32)
    out[:, :3] = p.astype(np.float32)
    return out


def rotation_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def make_T(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def reference_test_cloud():
    """generateCloud(), test/PointCloudRegistrationTest.cc:12-28: 30 x 50 grid, 0.5 spacing, z = sin x + cos y."""
    pts = []
    x = 0.0
    for _ in range(30):
        y = 0.0
        for _ in range(50):
            pts.append((np.float32(x), np.float32(y), np.float32(np.sin(x) + np.cos(y))))
            y += 0.5
        x += 0.5
    return _xyzw(np.array(pts, dtype=np.float32))


def reference_test_transform():
    """test/PointCloudRegistrationTest.cc:34-37: translation (2.5,0,0) then prerotate Rz(0.34)."""
    R = rotation_from_axis_angle([0, 0, 1], 0.34)
    return make_T(R, R @ np.array([2.5, 0.0, 0.0]))


def apply_T_like_pcl(cloud, T):
    """pcl::transformPointCloud(Affine3d): double math row by row, float32 store."""
    p = cloud[:, :3].astype(np.float64)
    out = cloud.copy()
    for r in range(3):
        out[:, r] = (T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1] + T[r, 2] * p[:, 2] + T[r, 3]).astype(np.float32)
    return out


def config1_plane_sphere(seed=1, n_plane=5000, n_sphere=5000, noise=0.01):
    """C1: plane z=0 on [-5,5]^2 plus a sphere r=1.5 at (0,0,1.5); source = R*target + t + noise.
    Returns (source, target, T_source_to_target)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    plane = np.zeros((n_plane, 3))
    plane[:, :2] = rng.uniform(-5, 5, size=(n_plane, 2))
    v = rng.normal(size=(n_sphere, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    sphere = 1.5 * v + np.array([0, 0, 1.5])
    target = np.concatenate([plane, sphere])
    R = rotation_from_axis_angle(np.array([1.0, 2.0, 3.0]) / np.sqrt(14.0), np.deg2rad(10.0))
    t = 0.1 * np.array([2.0, 1.0, 2.0]) / 3.0
    source = target @ R.T + t + rng.normal(scale=noise, size=target.shape)
    T_src_to_tgt = np.linalg.inv(make_T(R, t))
    return _xyzw(source), _xyzw(target), T_src_to_tgt


# -- LiDAR-like scene: ground plane z=-1.73 inside a 60 x 60 x 20 m box plus a few box obstacles ----------

_OBSTACLES = [  # axis-aligned boxes (cx, cy, half_x, half_y, height) standing on the ground
    (8.0, 3.0, 1.0, 2.2, 1.6), (-6.0, 9.0, 2.5, 1.0, 2.5), (14.0, -11.0, 1.5, 1.5, 3.0), (-15.0, -6.0, 1.0, 3.0, 2.0),
    (3.0, -17.0, 3.0, 1.0, 4.0), (-22.0, 14.0, 2.0, 2.0, 6.0), (21.0, 18.0, 2.5, 1.2, 2.2), (-3.0, 22.0, 4.0, 1.0, 3.0),
]
_GROUND_Z = -1.73
_HALF = 30.0
_CEIL = _GROUND_Z + 20.0


def _ray_box(o, inv, lo, hi):
    """Slab test, vectorised over rays (inv = 1 / direction, [N,3]); returns the entry distance (inf if missed).
    NaN slabs (0 * inf: a ray parallel to a face it starts on) are ignored like nanmax / nanmin would."""
    tn = tf = None
    for ax in range(3):
        t0 = (lo[ax] - o[ax]) * inv[:, ax]
        t1 = (hi[ax] - o[ax]) * inv[:, ax]
        a, b = np.minimum(t0, t1), np.maximum(t0, t1)
        tn = a if tn is None else np.fmax(tn, a)
        tf = b if tf is None else np.fmin(tf, b)
    hit = (tf >= tn) & (tn > 0.05)
    return np.where(hit, tn, np.inf)


def lidar_scan(rng, n_rings, n_az, pose=None, range_noise=0.02):
    """One sweep from sensor pose `pose` (4x4, sensor->world); points returned in the SENSOR frame."""
    pose = np.eye(4) if pose is None else pose
    elev = np.deg2rad(np.linspace(-24.8, 2.0, n_rings))
    az = np.linspace(0.0, 2 * np.pi, n_az, endpoint=False)
    e, a = np.meshgrid(elev, az, indexing="ij")
    d_s = np.stack([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)], axis=-1).reshape(-1, 3)
    R, o = pose[:3, :3], pose[:3, 3]
    d = d_s @ R.T
    t = np.full(len(d), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (_GROUND_Z - o[2]) / d[:, 2]
    t = np.where((tg > 0) & np.isfinite(tg), np.minimum(t, tg), t)
    # room: leave through one of the walls / the ceiling
    with np.errstate(divide="ignore", invalid="ignore"):
        for ax, lim in ((0, _HALF), (0, -_HALF), (1, _HALF), (1, -_HALF), (2, _CEIL)):
            tw = (lim - o[ax]) / d[:, ax]
            t = np.where((tw > 0) & np.isfinite(tw), np.minimum(t, tw), t)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        for cx, cy, hx, hy, h in _OBSTACLES:
            lo = np.array([cx - hx, cy - hy, _GROUND_Z])
            hi = np.array([cx + hx, cy + hy, _GROUND_Z + h])
            t = np.minimum(t, _ray_box(o, inv, lo, hi))
    t = t + rng.normal(scale=range_noise, size=t.shape)
    return d_s * t[:, None]


def lidar_pair(seed, n_rings, n_az, yaw_deg=2.0, trans=(0.5, 0.1, 0.0), outlier_frac=0.0, range_noise=0.02,
               random_motion=None):
    """Target = sweep from the origin; source = sweep of the same scene from a nearby pose, each in its own
    sensor frame.  Returns (source, target, T_source_to_target)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if random_motion is not None:
        max_deg, max_m = random_motion
        axis = rng.normal(size=3)
        axis[:2] *= 0.2  # mostly yaw, like a vehicle
        ang = np.deg2rad(rng.uniform(-max_deg, max_deg))
        R = rotation_from_axis_angle(axis, ang)
        tr = rng.uniform(-1, 1, size=3) * np.array([max_m, max_m * 0.3, max_m * 0.05])
    else:
        R = rotation_from_axis_angle([0, 0, 1], np.deg2rad(yaw_deg))
        tr = np.asarray(trans, dtype=np.float64)
    pose_b = make_T(R, tr)
    target = lidar_scan(rng, n_rings, n_az, np.eye(4), range_noise)
    source = lidar_scan(rng, n_rings, n_az, pose_b, range_noise)
    if outlier_frac > 0:
        for cloud in (target, source):
            n_out = int(round(outlier_frac * len(cloud)))
            sel = rng.choice(len(cloud), size=n_out, replace=False)
            lo = np.array([-_HALF, -_HALF, _GROUND_Z])
            hi = np.array([_HALF, _HALF, _CEIL])
            cloud[sel] = rng.uniform(lo, hi, size=(n_out, 3))
    return _xyzw(source), _xyzw(target), pose_b


def config2_lidar_outliers(seed=2, n_rings=64, n_az=1563):
    """C2: ~100k rays, 20% uniform outliers; run with -u -s 0.05 -t 0.05."""
    return lidar_pair(seed, n_rings, n_az, outlier_frac=0.2)


def config3_lidar_1m(seed=3, n_rings=128, n_az=7813):
    """C3: ~1M-point pair; run with -m 10 -r 0.5 -d 5."""
    return lidar_pair(seed, n_rings, n_az)


def config4_lidar_10m(seed=4, n_rings=320, n_az=31250):
    """C4: 10M-point pair, source-sharded across GPUs."""
    return lidar_pair(seed, n_rings, n_az)


def config5_pair(index, n_rings=64, n_az=1875):
    """C5: pair `index` of the 1024-pair batch (seeds 1000..2023), random motion <= 3 deg / <= 1 m."""
    return lidar_pair(1000 + index, n_rings, n_az, random_motion=(3.0, 1.0))


def pose_error(T_est, T_true):
    """(rotation error in rad, translation error in m) between two 4x4 transforms."""
    dR = T_est[:3, :3].T @ T_true[:3, :3]
    c = np.clip((np.trace(dR) - 1.0) / 2.0, -1.0, 1.0)
    return float(np.arccos(c)), float(np.linalg.norm(T_est[:3, 3] - T_true[:3, 3]))
