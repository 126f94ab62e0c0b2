"""ctypes binding of libppcr_cuda.so (include/ppcr.h) for the Python harness (tests, bench.py).

The product's host side is C++ (host/prob_point_cloud_registration.cc + the CLI); this module only lets Python
call the same C ABI.  It never falls back to a CPU path: loading fails loudly when the library is missing and
every call raises PpcrError when no B200 is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPCR_LIB_PATH") or os.path.join(_PKG, "csrc", "libppcr_cuda.so")  # override: tuning builds

EXPORTED_SYMBOLS = [
    "ppcr_last_error", "ppcr_version", "ppcr_default_params", "ppcr_default_options", "ppcr_create",
    "ppcr_create_ex", "ppcr_destroy", "ppcr_align", "ppcr_has_converged", "ppcr_history", "ppcr_increment_history",
    "ppcr_iteration_stats",
    "ppcr_filtered_source", "ppcr_filtered_target", "ppcr_association", "ppcr_get_stage_times", "ppcr_time_kernel",
    "ppcr_voxel_filter", "ppcr_time_voxel_filter", "ppcr_radius_search", "ppcr_weights_normal_eq", "ppcr_iteration_solve", "ppcr_transform", "ppcr_transform_ex",
    "ppcr_replay_metrics", "ppcr_closest_point_metrics",
    "ppcr_align_batch", "ppcr_align_batch_devices", "ppcr_host_alloc", "ppcr_host_free", "ppcr_shard_export", "ppcr_shard_connect", "ppcr_align_sharded",
]

SHARD_TOKEN_BYTES = 128


class PpcrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ppcr status {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """ppcr_params == ProbPointCloudRegistrationParams (params.hpp:5-18)."""

    _fields_ = [
        ("max_neighbours", C.c_int32),
        ("n_iter", C.c_int32),
        ("dof", C.c_double),
        ("radius", C.c_double),
        ("cost_drop_thresh", C.c_double),
        ("n_cost_drop_it", C.c_double),
        ("verbose", C.c_int32),
        ("summary", C.c_int32),
        ("initial_rotation", C.c_double * 4),
        ("initial_translation", C.c_double * 3),
        ("source_filter_size", C.c_double),
        ("target_filter_size", C.c_double),
    ]


class Options(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("input_on_device", C.c_int32),
        ("driver", C.c_int32),
        ("ticks_per_sync", C.c_int32),
        ("function_tolerance", C.c_double),
        ("leaf_capacity", C.c_int32),
        ("exact_weights", C.c_int32),
        ("stream", C.c_void_p),
        ("record_stage_times", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class IterStats(C.Structure):
    _fields_ = [
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("cost_drop", C.c_double),
        ("n_correspondences", C.c_int64),
        ("lm_iterations", C.c_int32),
        ("num_successful_steps", C.c_int32),
    ]


class StageTimes(C.Structure):
    _fields_ = [
        ("grid_build_ms", C.c_float), ("search_ms", C.c_float), ("eval_ms", C.c_float), ("controller_ms", C.c_float),
        ("transform_ms", C.c_float), ("voxel_ms", C.c_float),
        ("search_launches", C.c_int32), ("eval_launches", C.c_int32), ("controller_launches", C.c_int32),
        ("transform_launches", C.c_int32), ("total_launches", C.c_int32), ("ticks", C.c_int32),
        ("exchanges", C.c_int32), ("exchange_wait_ms", C.c_float),
    ]


class ClosestMetrics(C.Structure):
    """ppcr_closest_metrics: the seven closest-point helpers of utilities.hpp:28-234."""

    _fields_ = [
        ("average_closest_distance", C.c_double), ("sum_squared_error", C.c_double),
        ("robust_sum_squared_error", C.c_double), ("robust_sum_squared_error_factor", C.c_double),
        ("robust_averaged_sum_squared_error", C.c_double), ("median_closest_distance", C.c_double),
        ("robust_median_closest_distance", C.c_double), ("n_filtered", C.c_int64), ("n_filtered_factor", C.c_int64),
    ]


class PairDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("n_src", C.c_int64), ("tgt", C.c_void_p), ("n_tgt", C.c_int64)]


_lib = None


def lib():
    """Loads libppcr_cuda.so (no GPU needed to load; every compute call needs one)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -m probabilistic_point_clouds_registration_b200.build` "
                "(there is no CPU fallback)")
        # PPCR_CUDA_LIB: a differently tuned build of the same library (tools/tune_build.sh), tuning runs only
        L = C.CDLL(os.environ.get("PPCR_CUDA_LIB", LIB_PATH))
        if "PPCR_CUDA_LIB" in os.environ:  # an older build may lack the newest entry points: tolerate that in tuning runs
            class _Missing:
                pass

            class _Tolerant:
                def __getattr__(self, name):
                    try:
                        return getattr(L_real, name)
                    except AttributeError:
                        return _Missing()
            L_real, L = L, _Tolerant()
        vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.ppcr_last_error.restype = C.c_char_p
        L.ppcr_version.restype = C.c_char_p
        L.ppcr_default_params.argtypes = [C.POINTER(Params)]
        L.ppcr_default_options.argtypes = [C.POINTER(Options)]
        L.ppcr_create.argtypes = [vp, i64, vp, i64, C.POINTER(Params), C.POINTER(vp)]
        L.ppcr_create_ex.argtypes = [vp, i64, vp, i64, C.POINTER(Params), C.POINTER(Options), C.POINTER(vp)]
        L.ppcr_destroy.argtypes = [vp]
        L.ppcr_destroy.restype = None
        L.ppcr_align.argtypes = [vp]
        L.ppcr_has_converged.argtypes = [vp, C.POINTER(i32)]
        L.ppcr_history.argtypes = [vp, vp, C.POINTER(i32)]
        L.ppcr_increment_history.argtypes = [vp, vp, C.POINTER(i32)]
        L.ppcr_iteration_stats.argtypes = [vp, vp, C.POINTER(i32)]
        L.ppcr_filtered_source.argtypes = [vp, vp, C.POINTER(i64)]
        L.ppcr_filtered_target.argtypes = [vp, vp, C.POINTER(i64)]
        L.ppcr_association.argtypes = [vp, vp, vp, i64, i32]
        L.ppcr_get_stage_times.argtypes = [vp, C.POINTER(StageTimes)]
        L.ppcr_time_kernel.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(f64)]
        L.ppcr_voxel_filter.argtypes = [vp, i64, f64, vp, C.POINTER(i64)]
        L.ppcr_time_voxel_filter.argtypes = [vp, i64, f64, C.POINTER(Options), i32, C.POINTER(C.c_float), C.POINTER(f64), C.POINTER(i64)]
        L.ppcr_radius_search.argtypes = [vp, i64, vp, i64, f64, i32, i32, vp, vp, vp]
        L.ppcr_weights_normal_eq.argtypes = [vp, i64, vp, i64, vp, vp, i32, f64, i32, vp, vp, i32, vp, vp]
        L.ppcr_iteration_solve.argtypes = [vp, i64, vp, i64, vp, vp, i32, C.POINTER(Params), C.POINTER(Options), f64, vp, vp, vp]
        L.ppcr_transform.argtypes = [vp, i64, vp]
        L.ppcr_transform_ex.argtypes = [vp, i64, vp, C.POINTER(Options)]
        L.ppcr_replay_metrics.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
        L.ppcr_closest_point_metrics.argtypes = [vp, i64, vp, i64, f64, C.POINTER(Options), C.POINTER(ClosestMetrics), vp]
        L.ppcr_align_batch.argtypes = [C.POINTER(PairDesc), i32, C.POINTER(Params), C.POINTER(Options), i32, vp, vp, vp]
        L.ppcr_align_batch_devices.argtypes = [C.POINTER(PairDesc), i32, C.POINTER(Params), C.POINTER(Options), vp, i32, i32, vp, vp, vp]
        L.ppcr_host_alloc.argtypes = [C.c_size_t]
        L.ppcr_host_alloc.restype = vp
        L.ppcr_host_free.argtypes = [vp]
        L.ppcr_shard_export.argtypes = [vp, i32, i32, vp]
        L.ppcr_shard_connect.argtypes = [vp, vp]
        L.ppcr_align_sharded.argtypes = [vp, i64, vp, i64, C.POINTER(Params), C.POINTER(Options), vp, i32, vp, vp, vp]
        for name in EXPORTED_SYMBOLS:
            fn = getattr(L, name)
            if name not in ("ppcr_last_error", "ppcr_version", "ppcr_destroy", "ppcr_default_params", "ppcr_default_options",
                            "ppcr_host_alloc"):
                fn.restype = i32
        _lib = L
    return _lib


def _check(status):
    if status != 0:
        raise PpcrError(status, lib().ppcr_last_error().decode())


def _cloud(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, "clouds are [N,4] float32 (x,y,z,pad) like pcl::PointXYZ"
    return a


def make_params(max_neighbours=20, dof=5.0, radius=1.0, n_iter=1000, cost_drop_thresh=0.01, n_cost_drop_it=5.0,
                initial_rotation=(1.0, 0.0, 0.0, 0.0), initial_translation=(0.0, 0.0, 0.0), source_filter_size=0.0,
                target_filter_size=0.0) -> Params:
    """Struct defaults of params.hpp:6-17 (radius 1; the CLI's default radius is 3)."""
    p = Params()
    lib().ppcr_default_params(C.byref(p))
    p.max_neighbours = int(max_neighbours)
    p.dof = float(dof)
    p.radius = float(radius)
    p.n_iter = int(n_iter)
    p.cost_drop_thresh = float(cost_drop_thresh)
    p.n_cost_drop_it = float(n_cost_drop_it)
    p.initial_rotation[:] = [float(v) for v in initial_rotation]
    p.initial_translation[:] = [float(v) for v in initial_translation]
    p.source_filter_size = float(source_filter_size)
    p.target_filter_size = float(target_filter_size)
    return p


def make_options(device=0, input_on_device=False, driver=0, ticks_per_sync=0, function_tolerance=0.0, leaf_capacity=0,
                 exact_weights=False, stream=None, record_stage_times=False) -> Options:
    o = Options()
    lib().ppcr_default_options(C.byref(o))
    o.device = int(device)
    o.input_on_device = int(bool(input_on_device))
    o.driver = int(driver)
    o.ticks_per_sync = int(ticks_per_sync)
    o.function_tolerance = float(function_tolerance)
    o.leaf_capacity = int(leaf_capacity)
    o.exact_weights = int(bool(exact_weights))
    o.stream = C.c_void_p(stream) if stream else None
    o.record_stage_times = int(bool(record_stage_times))
    return o


class Registration:
    """Thin object over a ppcr_handle; mirrors ProbPointCloudRegistration's methods (registration.h:18-45)."""

    def __init__(self, source, target, params: Params, options: Options | None = None, n_source=None, n_target=None):
        """source/target: [N,4] float32 numpy arrays, or raw device pointers (ints) with n_source/n_target when
        options.input_on_device is set."""
        self._h = C.c_void_p()
        self.params = params
        if options is not None and options.input_on_device:
            src_ptr, n_src, tgt_ptr, n_tgt = int(source), int(n_source), int(target), int(n_target)
        else:
            self._src = _cloud(source)
            self._tgt = _cloud(target)
            src_ptr, n_src = self._src.ctypes.data, len(self._src)
            tgt_ptr, n_tgt = self._tgt.ctypes.data, len(self._tgt)
        _check(lib().ppcr_create_ex(src_ptr, n_src, tgt_ptr, n_tgt, C.byref(params),
                                    C.byref(options) if options is not None else None, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ppcr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def align(self):
        _check(lib().ppcr_align(self._h))

    def has_converged(self) -> bool:
        out = C.c_int32(0)
        _check(lib().ppcr_has_converged(self._h, C.byref(out)))
        return bool(out.value)

    def transformation_history(self) -> np.ndarray:
        n = C.c_int32(0)
        _check(lib().ppcr_history(self._h, None, C.byref(n)))
        hist = np.zeros((max(n.value, 1), 16))
        cap = C.c_int32(n.value)
        _check(lib().ppcr_history(self._h, hist.ctypes.data, C.byref(cap)))
        return hist[:n.value].reshape(n.value, 4, 4)

    def increment_history(self) -> np.ndarray:
        n = C.c_int32(0)
        _check(lib().ppcr_increment_history(self._h, None, C.byref(n)))
        hist = np.zeros((max(n.value, 1), 16))
        cap = C.c_int32(n.value)
        _check(lib().ppcr_increment_history(self._h, hist.ctypes.data, C.byref(cap)))
        return hist[:n.value].reshape(n.value, 4, 4)

    def replay_metrics(self, cloud, ground_truth=None, first=0, count=None):
        """Per-iteration diagnostics of the reference on the device (registration.cc:110-122): moves `cloud` ([N,4] float32)
        by the increments of outer iterations [first, first + count) and returns (moved cloud, mean distance to the ground
        truth per iteration, mean distance to the previous position per iteration)."""
        cloud = np.ascontiguousarray(cloud, dtype=np.float32).copy()
        gt = None if ground_truth is None else np.ascontiguousarray(ground_truth, dtype=np.float32)
        if count is None:
            count = len(self.iteration_stats()) - first
        mse_gt, mse_prev = np.zeros(max(count, 1)), np.zeros(max(count, 1))
        _check(lib().ppcr_replay_metrics(self._h, cloud.ctypes.data, None if gt is None else gt.ctypes.data, len(cloud),
                                         int(first), int(count), mse_gt.ctypes.data, mse_prev.ctypes.data))
        return cloud, mse_gt[:count], mse_prev[:count]

    def transformation(self) -> np.ndarray:
        h = self.transformation_history()
        if len(h) == 0:
            raise IndexError("transformation(): no outer iteration has run (the reference reads .back() of an empty vector)")
        return h[-1]

    def iteration_stats(self):
        n = C.c_int32(0)
        _check(lib().ppcr_iteration_stats(self._h, None, C.byref(n)))
        arr = (IterStats * max(n.value, 1))()
        cap = C.c_int32(n.value)
        _check(lib().ppcr_iteration_stats(self._h, arr, C.byref(cap)))
        return [dict(initial_cost=s.initial_cost, final_cost=s.final_cost, cost_drop=s.cost_drop,
                     n_correspondences=s.n_correspondences, lm_iterations=s.lm_iterations,
                     num_successful_steps=s.num_successful_steps) for s in arr[:n.value]]

    def _download(self, fn):
        n = C.c_int64(0)
        _check(fn(self._h, None, C.byref(n)))
        out = np.zeros((max(n.value, 1), 4), dtype=np.float32)
        cap = C.c_int64(n.value)
        _check(fn(self._h, out.ctypes.data, C.byref(cap)))
        return out[:n.value]

    def filtered_source(self) -> np.ndarray:
        return self._download(lib().ppcr_filtered_source)

    def filtered_target(self) -> np.ndarray:
        return self._download(lib().ppcr_filtered_target)

    def association(self):
        n = len(self.filtered_source())
        m = self.params.max_neighbours
        idx = np.full((n, m), -1, dtype=np.int32)
        cnt = np.zeros(n, dtype=np.int32)
        _check(lib().ppcr_association(self._h, idx.ctypes.data, cnt.ctypes.data, n, m))
        return idx, cnt

    def stage_times(self) -> StageTimes:
        st = StageTimes()
        _check(lib().ppcr_get_stage_times(self._h, C.byref(st)))
        return st

    def time_kernel(self, which: int, reps: int = 10, flush_l2: bool = True):
        """which: 0 search, 1 eval (weights + normal equations), 2 transform, 3 grid build.
        Returns (average milliseconds per launch, algorithmic bytes per launch)."""
        ms = C.c_float(0)
        by = C.c_double(0)
        _check(lib().ppcr_time_kernel(self._h, int(which), int(reps), int(bool(flush_l2)), C.byref(ms), C.byref(by)))
        return ms.value, by.value


def voxel_filter(cloud, leaf):
    cloud = _cloud(cloud)
    out = np.zeros_like(cloud)
    n = C.c_int64(0)
    _check(lib().ppcr_voxel_filter(cloud.ctypes.data, len(cloud), float(leaf), out.ctypes.data, C.byref(n)))
    return out[:n.value].copy()


def voxel_filter_timed(cloud_ptr, n, leaf, device=0, reps=3):
    """Average device time of the voxel filter on a DEVICE-resident cloud; dict for bench.py's roofline.kernels."""
    opt = make_options(device=device, input_on_device=True)
    ms, by, k = C.c_float(0), C.c_double(0), C.c_int64(0)
    _check(lib().ppcr_time_voxel_filter(int(cloud_ptr), int(n), float(leaf), C.byref(opt), int(reps), C.byref(ms), C.byref(by),
                                        C.byref(k)))
    if k.value < 0:
        return None
    return {"avg_ms": ms.value, "algorithmic_bytes": by.value, "gbs": by.value / (ms.value * 1e-3) / 1e9, "leaf": leaf,
            "n_in": int(n), "n_out": int(k.value)}


def radius_search(src, tgt, radius, max_nn, leaf_capacity=0):
    src, tgt = _cloud(src), _cloud(tgt)
    cols = max(int(max_nn), 1)  # invalid max_nn values are rejected by the library, not by numpy
    idx = np.full((len(src), cols), -1, dtype=np.int32)
    d2 = np.zeros((len(src), cols), dtype=np.float32)
    cnt = np.zeros(len(src), dtype=np.int32)
    _check(lib().ppcr_radius_search(src.ctypes.data, len(src), tgt.ctypes.data, len(tgt), float(radius), int(max_nn),
                                    int(leaf_capacity), idx.ctypes.data, d2.ctypes.data, cnt.ctypes.data))
    return idx, d2, cnt


def weights_normal_eq(src, tgt, idx, count, dof, pose_w, pose_e, fast_weights=False, want_weights=True, dimension=3):
    src, tgt = _cloud(src), _cloud(tgt)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.int32)
    max_nn = idx.shape[1]
    pw = np.ascontiguousarray(pose_w, dtype=np.float64)
    pe = np.ascontiguousarray(pose_e, dtype=np.float64)
    w = np.zeros(idx.shape, dtype=np.float64) if want_weights else None
    ne = np.zeros(36)
    _check(lib().ppcr_weights_normal_eq(src.ctypes.data, len(src), tgt.ctypes.data, len(tgt), idx.ctypes.data,
                                        count.ctypes.data, max_nn, float(dof), int(dimension), pw.ctypes.data, pe.ctypes.data,
                                        int(bool(fast_weights)), w.ctypes.data if w is not None else None,
                                        ne.ctypes.data))
    return w, ne


def iteration_solve(src, tgt, idx, count, params: Params, function_tolerance=1e-5, options: Options | None = None):
    src, tgt = _cloud(src), _cloud(tgt)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    if idx.ndim == 1:
        idx = idx.reshape(-1, 1)
    count = np.ascontiguousarray(count, dtype=np.int32)
    pose = np.zeros(7)
    T = np.zeros(16)
    st = IterStats()
    _check(lib().ppcr_iteration_solve(src.ctypes.data, len(src), tgt.ctypes.data, len(tgt), idx.ctypes.data,
                                      count.ctypes.data, idx.shape[1], C.byref(params),
                                      C.byref(options) if options is not None else None, float(function_tolerance),
                                      pose.ctypes.data, T.ctypes.data, C.byref(st)))
    return pose, T.reshape(4, 4), dict(initial_cost=st.initial_cost, final_cost=st.final_cost,
                                       lm_iterations=st.lm_iterations, num_successful_steps=st.num_successful_steps,
                                       n_correspondences=st.n_correspondences)


def transform(cloud, T, options: Options | None = None, n_points=None):
    """pcl::transformPointCloud.  With options.input_on_device `cloud` is a device pointer (moved in place, n_points given)."""
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    if options is not None and options.input_on_device:
        _check(lib().ppcr_transform_ex(int(cloud), int(n_points), T.ctypes.data, C.byref(options)))
        return None
    out = _cloud(cloud).copy()
    if options is None:
        _check(lib().ppcr_transform(out.ctypes.data, len(out), T.ctypes.data))
    else:
        _check(lib().ppcr_transform_ex(out.ctypes.data, len(out), T.ctypes.data, C.byref(options)))
    return out


def closest_point_metrics(cloud1, cloud2, factor=3.0, options: Options | None = None, want_distances=True):
    """utilities.hpp:28-234: dict of the seven helpers' values (+ window counts), and the squared 1-NN distance of every
    cloud1 point in cloud2 (cloud1's order)."""
    a, b = _cloud(cloud1), _cloud(cloud2)
    out = ClosestMetrics()
    d2 = np.zeros(max(len(a), 1), dtype=np.float32) if want_distances else None
    _check(lib().ppcr_closest_point_metrics(a.ctypes.data, len(a), b.ctypes.data, len(b), float(factor),
                                            C.byref(options) if options is not None else None, C.byref(out),
                                            d2.ctypes.data if d2 is not None else None))
    return {name: getattr(out, name) for name, _ in ClosestMetrics._fields_}, (d2[:len(a)] if d2 is not None else None)


def align_batch(pairs, params: Params, options: Options | None = None, slots=0, devices=None):
    """pairs: list of (source, target) numpy clouds (or (src_ptr, n_src, tgt_ptr, n_tgt) device tuples when
    options.input_on_device).  devices: list of device ordinals (ppcr_align_batch_devices: `slots` lanes on each).
    Returns (T [n,4,4], n_outer [n], correspondences [n])."""
    n = len(pairs)
    descs = (PairDesc * max(n, 1))()
    keep = []
    for i, pr in enumerate(pairs):
        if options is not None and options.input_on_device:
            descs[i].src, descs[i].n_src, descs[i].tgt, descs[i].n_tgt = pr
        else:
            s, t = _cloud(pr[0]), _cloud(pr[1])
            keep.append((s, t))
            descs[i].src, descs[i].n_src, descs[i].tgt, descs[i].n_tgt = s.ctypes.data, len(s), t.ctypes.data, len(t)
    T = np.zeros((max(n, 1), 16))
    n_outer = np.zeros(max(n, 1), dtype=np.int32)
    corr = np.zeros(max(n, 1), dtype=np.int64)
    popt = C.byref(options) if options is not None else None
    if devices is None:
        _check(lib().ppcr_align_batch(descs, n, C.byref(params), popt, int(slots), T.ctypes.data, n_outer.ctypes.data,
                                      corr.ctypes.data))
    else:
        ids = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        _check(lib().ppcr_align_batch_devices(descs, n, C.byref(params), popt, ids, len(devices), int(slots), T.ctypes.data,
                                              n_outer.ctypes.data, corr.ctypes.data))
    return T[:n].reshape(n, 4, 4), n_outer[:n], corr[:n]


def align_sharded(source, target, params: Params, devices, options: Options | None = None, max_history=4096):
    """ppcr_align_sharded: one pair over the GPUs `devices` of this process.  Returns (history [n,4,4], correspondences)."""
    s, t = _cloud(source), _cloud(target)
    ids = (C.c_int32 * len(devices))(*[int(d) for d in devices])
    hist = np.zeros((max_history, 16))
    n = C.c_int32(max_history)
    corr = C.c_int64(0)
    _check(lib().ppcr_align_sharded(s.ctypes.data, len(s), t.ctypes.data, len(t), C.byref(params),
                                    C.byref(options) if options is not None else None, ids, len(devices), hist.ctypes.data,
                                    C.byref(n), C.byref(corr)))
    k = min(n.value, max_history)
    return hist[:k].reshape(k, 4, 4), int(corr.value)
