"""ctypes loader for the CPU oracle (oracle/libppcr_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package never does (tests/test_no_oracle_in_product.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libppcr_oracle.so")
_lib = None


class OracleParams(C.Structure):
    """Mirror of oracle_params (ppcr_oracle.h) == ProbPointCloudRegistrationParams (params.hpp:5-18)."""

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


class SolverOptions(C.Structure):
    _fields_ = [
        ("function_tolerance", C.c_double),
        ("max_num_iterations", C.c_int32),
        ("inner_kind", C.c_int32),
        ("num_threads", C.c_int32),
    ]


class SolveSummary(C.Structure):
    _fields_ = [
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("num_iterations", C.c_int32),
        ("num_successful_steps", C.c_int32),
        ("termination", C.c_int32),
        ("num_nonmonotonic_steps", C.c_int32),
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


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ppcr_oracle.cpp")
    hdr = os.path.join(_HERE, "ppcr_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libppcr_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()  # (re)compiles only when the library is missing or older than its sources
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        lp = C.POINTER(C.c_int64)
        L.oracle_radius_search.restype = C.c_int64
        L.oracle_radius_search.argtypes = [fp, C.c_int64, fp, C.c_int64, C.c_double, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, ip, fp, ip]
        L.oracle_update_weights.restype = None
        L.oracle_update_weights.argtypes = [C.c_int64, lp, dp, C.c_double, C.c_int32, dp]
        L.oracle_callback_weights.restype = None
        L.oracle_callback_weights.argtypes = [fp, fp, C.c_int64, lp, ip, dp, dp, C.c_double, dp, dp]
        L.oracle_iteration_solve.restype = C.c_int32
        L.oracle_iteration_solve.argtypes = [fp, C.c_int64, fp, C.c_int64, lp, ip, C.POINTER(OracleParams),
                                             C.POINTER(SolverOptions), dp, dp, dp, C.POINTER(SolveSummary)]
        L.oracle_transform.restype = None
        L.oracle_transform.argtypes = [fp, C.c_int64, dp]
        L.oracle_voxel_grid.restype = C.c_int64
        L.oracle_voxel_grid.argtypes = [fp, C.c_int64, C.c_double, fp]
        L.oracle_calculate_mse.restype = C.c_double
        L.oracle_calculate_mse.argtypes = [fp, fp, C.c_int64]
        L.oracle_align.restype = C.c_int32
        L.oracle_align.argtypes = [fp, C.c_int64, fp, C.c_int64, C.POINTER(OracleParams), C.POINTER(SolverOptions),
                                   C.c_int32, dp, C.POINTER(IterStats), C.c_int32, fp, lp, lp]
        L.oracle_max_threads.restype = C.c_int32
        L.oracle_closest_metrics.restype = C.c_int32
        L.oracle_closest_metrics.argtypes = [fp, C.c_int64, fp, C.c_int64, C.c_double, dp, fp]
        L.oracle_normal_eq.restype = None
        L.oracle_normal_eq.argtypes = [fp, fp, C.c_int64, lp, ip, C.c_double, dp, dp, dp]
        L.oracle_nonmonotonic_steps.restype = C.c_int64
        L.oracle_nonmonotonic_steps.argtypes = [C.c_int32]
        _lib = L
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, "clouds are [N,4] float32 (x,y,z,pad) like pcl::PointXYZ"
    return a


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def make_params(max_neighbours=20, dof=5.0, radius=1.0, n_iter=1000, cost_drop_thresh=0.01, n_cost_drop_it=5.0,
                initial_rotation=(1.0, 0.0, 0.0, 0.0), initial_translation=(0.0, 0.0, 0.0),
                source_filter_size=0.0, target_filter_size=0.0) -> OracleParams:
    """Struct defaults of params.hpp:6-17 (radius 1, NOT the CLI's 3)."""
    p = OracleParams()
    p.max_neighbours = int(max_neighbours)
    p.n_iter = int(n_iter)
    p.dof = float(dof)
    p.radius = float(radius)
    p.cost_drop_thresh = float(cost_drop_thresh)
    p.n_cost_drop_it = float(n_cost_drop_it)
    p.verbose = 0
    p.summary = 0
    p.initial_rotation[:] = [float(v) for v in initial_rotation]
    p.initial_translation[:] = [float(v) for v in initial_translation]
    p.source_filter_size = float(source_filter_size)
    p.target_filter_size = float(target_filter_size)
    return p


def make_options(function_tolerance=1e-5, max_num_iterations=2**31 - 1, inner_kind=0, num_threads=0) -> SolverOptions:
    o = SolverOptions()
    o.function_tolerance = float(function_tolerance)
    o.max_num_iterations = int(max_num_iterations)
    o.inner_kind = int(inner_kind)
    o.num_threads = int(num_threads)
    return o


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def nonmonotonic_steps(reset=False) -> int:
    """Accepted non-monotonic LM steps over every solve since the last reset."""
    return int(lib().oracle_nonmonotonic_steps(int(bool(reset))))


def radius_search(src, tgt, radius, max_nn, use_grid=False, num_threads=0, cap=None):
    src, tgt = _f32(src), _f32(tgt)
    ns, nt = len(src), len(tgt)
    u = max_nn & 0xFFFFFFFF
    limit = nt if (u == 0 or u > nt) else u
    cap = int(limit if cap is None else cap)
    idx = np.full((ns, max(cap, 1)), -1, dtype=np.int32)
    d2 = np.zeros((ns, max(cap, 1)), dtype=np.float32)
    cnt = np.zeros(ns, dtype=np.int32)
    total = lib().oracle_radius_search(_ptr(src, C.c_float), ns, _ptr(tgt, C.c_float), nt, float(radius), int(max_nn),
                                       max(cap, 1), int(bool(use_grid)), int(num_threads), _ptr(idx, C.c_int32),
                                       _ptr(d2, C.c_float), _ptr(cnt, C.c_int32))
    return idx, d2, cnt, int(total)


def update_weights(row_ptr, squared_errors, dof, dimension):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    se = np.ascontiguousarray(squared_errors, dtype=np.float64)
    out = np.zeros_like(se)
    lib().oracle_update_weights(len(row_ptr) - 1, _ptr(row_ptr, C.c_int64), _ptr(se, C.c_double), float(dof),
                                int(dimension), _ptr(out, C.c_double))
    return out


def callback_weights(src, tgt, row_ptr, col_idx, rotation, translation, dof):
    src, tgt = _f32(src), _f32(tgt)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    rot = np.ascontiguousarray(rotation, dtype=np.float64)
    tr = np.ascontiguousarray(translation, dtype=np.float64)
    se = np.zeros(len(col_idx), dtype=np.float64)
    w = np.zeros(len(col_idx), dtype=np.float64)
    lib().oracle_callback_weights(_ptr(src, C.c_float), _ptr(tgt, C.c_float), len(row_ptr) - 1, _ptr(row_ptr, C.c_int64),
                                  _ptr(col_idx, C.c_int32), _ptr(rot, C.c_double), _ptr(tr, C.c_double), float(dof),
                                  _ptr(se, C.c_double), _ptr(w, C.c_double))
    return se, w


@dataclass
class SolveResult:
    rotation: np.ndarray
    translation: np.ndarray
    T: np.ndarray
    initial_cost: float
    final_cost: float
    num_iterations: int
    num_successful_steps: int
    termination: int


def iteration_solve(src, tgt, row_ptr, col_idx, params: OracleParams, options: SolverOptions) -> SolveResult:
    src, tgt = _f32(src), _f32(tgt)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    rot = np.zeros(4)
    tr = np.zeros(3)
    T = np.zeros(16)
    s = SolveSummary()
    lib().oracle_iteration_solve(_ptr(src, C.c_float), len(src), _ptr(tgt, C.c_float), len(tgt), _ptr(row_ptr, C.c_int64),
                                 _ptr(col_idx, C.c_int32), C.byref(params), C.byref(options), _ptr(rot, C.c_double),
                                 _ptr(tr, C.c_double), _ptr(T, C.c_double), C.byref(s))
    return SolveResult(rot, tr, T.reshape(4, 4), s.initial_cost, s.final_cost, s.num_iterations, s.num_successful_steps,
                       s.termination)


def transform(cloud, T):
    out = _f32(cloud).copy()
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    lib().oracle_transform(_ptr(out, C.c_float), len(out), _ptr(T, C.c_double))
    return out


def voxel_grid(cloud, leaf):
    cloud = _f32(cloud)
    out = np.zeros_like(cloud)
    n = lib().oracle_voxel_grid(_ptr(cloud, C.c_float), len(cloud), float(leaf), _ptr(out, C.c_float))
    if n < 0:
        return out, True
    return out[:n].copy(), False


CLOSEST_METRIC_NAMES = ("average_closest_distance", "sum_squared_error", "robust_sum_squared_error",
                        "robust_sum_squared_error_factor", "robust_averaged_sum_squared_error", "median_closest_distance",
                        "robust_median_closest_distance", "n_filtered", "n_filtered_factor")


def closest_metrics(cloud1, cloud2, factor=3.0):
    """utilities.hpp:28-234 in one call: dict of the seven helpers' return values (+ the two window counts) and the
    per-point squared 1-NN distances."""
    a, b = _f32(cloud1), _f32(cloud2)
    out = np.zeros(9)
    d2 = np.zeros(len(a), dtype=np.float32)
    rc = lib().oracle_closest_metrics(_ptr(a, C.c_float), len(a), _ptr(b, C.c_float), len(b), float(factor),
                                      _ptr(out, C.c_double), _ptr(d2, C.c_float))
    if rc != 0:
        raise ValueError("empty cloud")
    return dict(zip(CLOSEST_METRIC_NAMES, out.tolist())), d2


def normal_eq(src, tgt, row_ptr, col_idx, dof, pose_w, pose_e):
    """J^T W J (upper triangle, 28), J^T W r (7), cost of one evaluation at pose_e with weights refreshed at pose_w."""
    src, tgt = _f32(src), _f32(tgt)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    xw = np.ascontiguousarray(pose_w, dtype=np.float64)
    xe = np.ascontiguousarray(pose_e, dtype=np.float64)
    out = np.zeros(36)
    lib().oracle_normal_eq(_ptr(src, C.c_float), _ptr(tgt, C.c_float), len(row_ptr) - 1, _ptr(row_ptr, C.c_int64),
                           _ptr(col_idx, C.c_int32), float(dof), _ptr(xw, C.c_double), _ptr(xe, C.c_double),
                           _ptr(out, C.c_double))
    return out


def calculate_mse(a, b):
    a, b = _f32(a), _f32(b)
    assert len(a) == len(b)
    return float(lib().oracle_calculate_mse(_ptr(a, C.c_float), _ptr(b, C.c_float), len(a)))


@dataclass
class AlignResult:
    history: np.ndarray  # [n_outer, 4, 4]
    stats: list = field(default_factory=list)
    filtered_source: np.ndarray | None = None
    n_filtered_src: int = 0
    n_filtered_tgt: int = 0

    @property
    def n_outer(self):
        return len(self.history)

    @property
    def transformation(self):
        return self.history[-1]


def align(src, tgt, params: OracleParams, options: SolverOptions, use_grid=True, max_hist=None) -> AlignResult:
    src, tgt = _f32(src), _f32(tgt)
    max_hist = int(params.n_iter if max_hist is None else max_hist)
    max_hist = max(1, min(max_hist, 100000))
    hist = np.zeros((max_hist, 16))
    stats = (IterStats * max_hist)()
    out_src = np.zeros_like(src)
    nfs = C.c_int64(0)
    nft = C.c_int64(0)
    n = lib().oracle_align(_ptr(src, C.c_float), len(src), _ptr(tgt, C.c_float), len(tgt), C.byref(params),
                           C.byref(options), int(bool(use_grid)), _ptr(hist, C.c_double), stats, max_hist,
                           _ptr(out_src, C.c_float), C.byref(nfs), C.byref(nft))
    n_rec = min(n, max_hist)
    st = [dict(initial_cost=stats[i].initial_cost, final_cost=stats[i].final_cost, cost_drop=stats[i].cost_drop,
               n_correspondences=stats[i].n_correspondences, lm_iterations=stats[i].lm_iterations,
               num_successful_steps=stats[i].num_successful_steps) for i in range(n_rec)]
    res = AlignResult(hist[:n_rec].reshape(n_rec, 4, 4).copy(), st, out_src[:nfs.value].copy(), nfs.value, nft.value)
    res.n_total = n
    return res


def weights_closed_form(sq_err_rows, dof, dimension):
    """numpy mirror of probabilistic_weights.hpp:48-105 for ONE row."""
    r2 = np.asarray(sq_err_rows, dtype=np.float64)
    if np.isinf(dof):
        lp = -r2 / 2.0
        p = np.exp(lp - lp.max())
        return p / p.sum()
    lp = -(dof + dimension) / 2.0 * np.log1p(r2 / dof)
    p = np.exp(lp - lp.max())
    return p / p.sum() * (dof + dimension) / (dof + r2)
