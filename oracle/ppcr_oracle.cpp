/*
 * ppcr_oracle.cpp -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see ppcr_oracle.h for who may load it and for the parity-pinning status of each part).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Third-party arithmetic that is absent from /root/reference is restated from the published
 * behaviour of those libraries and is kept in clearly marked sections:
 *   [FLANN/PCL]  pcl::KdTreeFLANN::radiusSearch, pcl::VoxelGrid, pcl::transformPointCloud
 *   [CERES]      ceres::Solve (trust-region Levenberg-Marquardt, DENSE_QR), AutoDiff, ScaledLoss
 *   [EIGEN]      Quaternion::normalize / toRotationMatrix, Affine3d product
 *
 * Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).  -ffp-contract=off is
 * load-bearing: FLANN's L2_Simple distance and PCL's transform have no FMA on a default x86-64 build.
 */
#include "ppcr_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kNumParams = 7;  // rotation_[4] + translation_[3]  (iteration.hpp:42-44)

int resolve_threads(int requested)
{
#ifdef _OPENMP
    // <= 0: OpenMP's own default (honours OMP_NUM_THREADS); an explicit request is only capped by the processors there are,
    // so that a caller can ask for every core even under a launcher that exports OMP_NUM_THREADS=1 (torchrun does)
    if (requested <= 0) return omp_get_max_threads();
    const int procs = omp_get_num_procs();
    return requested > procs ? procs : requested;
#else
    (void)requested;
    return 1;
#endif
}

// ---------------------------------------------------------------------------------------------
// [FLANN/PCL] radius search
// ---------------------------------------------------------------------------------------------

// FLANN L2_Simple<float>: accumulate (a-b)^2 over x, y, z in float32, in that order.
inline float dist2_f32(const float* a, const float* b)
{
    float dx = a[0] - b[0];
    float dy = a[1] - b[1];
    float dz = a[2] - b[2];
    float acc = dx * dx;
    acc = acc + dy * dy;
    acc = acc + dz * dz;
    return acc;
}

struct DistIndex {
    float d2;
    int32_t idx;
};
inline bool di_less(const DistIndex& a, const DistIndex& b)
{
    return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
}

// Bounded result set: the members are the points with d2 < r2 (strict, FLANN), and when more than `capacity` qualify the
// capacity smallest under the lexicographic order (d2, index) -- the tie rule BASELINE.json's north_star states
// ("distance, then index tiebreak"), which makes the result a pure function of the two clouds.  FLANN's own
// KNNRadiusResultSet drops a candidate whose d2 EQUALS the current worst whatever its index, so on an exact float tie at
// the m-th boundary its answer depends on the order its tree happens to visit the points in; there is no reference-held
// vector for that case (radius search unpinned), and an order-dependent rule cannot be restated without the
// tree.  Both the brute-force and the grid path below, and the CUDA kernels, use the pure rule.
struct ResultSet {
    std::vector<DistIndex> heap;
    size_t capacity;
    float r2;
    void reset(size_t cap, float r2f)
    {
        heap.clear();
        capacity = cap;
        r2 = r2f;
    }
    inline void offer(float d2, int32_t idx)
    {
        if (!(d2 < r2)) return;
        const DistIndex cand{d2, idx};
        if (heap.size() == capacity) {
            if (!di_less(cand, heap.front())) return;
            std::pop_heap(heap.begin(), heap.end(), di_less);
            heap.back() = cand;
        } else {
            heap.push_back(cand);
        }
        std::push_heap(heap.begin(), heap.end(), di_less);
    }
    // largest kept d2 once the set is full (nothing farther can enter), the squared radius before that
    float worst_d2() const { return heap.size() == capacity ? heap.front().d2 : r2; }
    bool full() const { return heap.size() == capacity; }
    void sorted(std::vector<DistIndex>& out)
    {
        out = heap;
        std::sort(out.begin(), out.end(), di_less);
    }
};

size_t effective_limit(int32_t max_nn, int64_t n_tgt)
{
    // pcl::KdTreeFLANN::radiusSearch: max_nn is unsigned; 0 or > N_t means "all points".
    uint32_t u = static_cast<uint32_t>(max_nn);
    if (u == 0 || static_cast<uint64_t>(u) > static_cast<uint64_t>(n_tgt)) return static_cast<size_t>(n_tgt);
    return u;
}

struct CpuGrid {
    float origin[3];
    float inv_h;
    float h;
    int dims[3];
    std::vector<int32_t> cell_start;  // ncells + 1
    std::vector<int32_t> order;       // target indices grouped by cell, ascending index within a cell
    inline int coord(float v, int a) const
    {
        int c = static_cast<int>(std::floor((v - origin[a]) * inv_h));
        return c;
    }
};

uint64_t count_occupied(const float* pts, int64_t n, const float* lo, float inv_h)
{
    std::vector<uint64_t> keys(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
        uint64_t cx = static_cast<uint64_t>(std::floor((pts[4 * i + 0] - lo[0]) * inv_h));
        uint64_t cy = static_cast<uint64_t>(std::floor((pts[4 * i + 1] - lo[1]) * inv_h));
        uint64_t cz = static_cast<uint64_t>(std::floor((pts[4 * i + 2] - lo[2]) * inv_h));
        keys[i] = (cx << 42) | (cy << 21) | cz;
    }
    std::sort(keys.begin(), keys.end());
    return static_cast<uint64_t>(std::unique(keys.begin(), keys.end()) - keys.begin());
}

void build_grid(const float* tgt, int64_t n, double radius, size_t limit, CpuGrid& g)
{
    float lo[3] = {tgt[0], tgt[1], tgt[2]}, hi[3] = {tgt[0], tgt[1], tgt[2]};
    for (int64_t i = 1; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], tgt[4 * i + a]);
            hi[a] = std::max(hi[a], tgt[4 * i + a]);
        }
    // Cell edge: start at the radius and halve while cells stay well populated relative to the
    // result-set size; the search below is exact for any edge, the edge only changes its cost.
    double h = radius;
    double max_ext = std::max({double(hi[0] - lo[0]), double(hi[1] - lo[1]), double(hi[2] - lo[2]), 1e-3});
    if (h > max_ext) h = max_ext;
    double want = std::max<double>(1.0, std::min<double>(limit, 64) / 3.0);
    for (int it = 0; it < 8; ++it) {
        double trial = h * 0.5;
        if (max_ext / trial > 1000.0) break;
        uint64_t occ = count_occupied(tgt, n, lo, static_cast<float>(1.0 / trial));
        if (static_cast<double>(n) / static_cast<double>(occ) < want) break;
        h = trial;
    }
    g.h = static_cast<float>(h);
    g.inv_h = 1.0f / g.h;
    int64_t ncells = 1;
    for (int a = 0; a < 3; ++a) {
        g.origin[a] = lo[a];
        g.dims[a] = static_cast<int>(std::floor((hi[a] - lo[a]) * g.inv_h)) + 1;
        ncells *= g.dims[a];
    }
    g.cell_start.assign(static_cast<size_t>(ncells) + 1, 0);
    std::vector<int32_t> cell_of(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
        int cx = std::min(std::max(g.coord(tgt[4 * i + 0], 0), 0), g.dims[0] - 1);
        int cy = std::min(std::max(g.coord(tgt[4 * i + 1], 1), 0), g.dims[1] - 1);
        int cz = std::min(std::max(g.coord(tgt[4 * i + 2], 2), 0), g.dims[2] - 1);
        int32_t c = (cz * g.dims[1] + cy) * g.dims[0] + cx;
        cell_of[i] = c;
        g.cell_start[static_cast<size_t>(c) + 1]++;
    }
    for (size_t c = 0; c < static_cast<size_t>(ncells); ++c) g.cell_start[c + 1] += g.cell_start[c];
    g.order.resize(static_cast<size_t>(n));
    std::vector<int32_t> cursor(g.cell_start.begin(), g.cell_start.end() - 1);
    for (int64_t i = 0; i < n; ++i) g.order[cursor[cell_of[i]]++] = static_cast<int32_t>(i);
}

// Exact bounded radius search on the grid: Chebyshev shells around the query cell, stopping when the set
// is full and its worst distance is already inside the region fully covered by the shells scanned so far.
void grid_query(const CpuGrid& g, const float* tgt, const float* q, float /*r2f*/, double radius, ResultSet& rs)
{
    const float rpad = static_cast<float>(radius * (1.0 + 1e-5)) + 1e-30f;
    int lo[3], hi[3], qc[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = g.coord(std::nextafter(q[a] - rpad, -INFINITY), a);
        hi[a] = g.coord(std::nextafter(q[a] + rpad, INFINITY), a);
        qc[a] = g.coord(q[a], a);
        lo[a] = std::max(lo[a], 0);
        hi[a] = std::min(hi[a], g.dims[a] - 1);
    }
    if (lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]) return;
    int smax = 0;
    for (int a = 0; a < 3; ++a) smax = std::max({smax, qc[a] - lo[a], hi[a] - qc[a]});
    for (int s = 0; s <= smax; ++s) {
        for (int cz = std::max(lo[2], qc[2] - s); cz <= std::min(hi[2], qc[2] + s); ++cz)
            for (int cy = std::max(lo[1], qc[1] - s); cy <= std::min(hi[1], qc[1] + s); ++cy) {
                bool face = (std::abs(cz - qc[2]) == s) || (std::abs(cy - qc[1]) == s);
                for (int cx = std::max(lo[0], qc[0] - s); cx <= std::min(hi[0], qc[0] + s); ++cx) {
                    if (!face && std::abs(cx - qc[0]) != s) continue;
                    size_t c = (static_cast<size_t>(cz) * g.dims[1] + cy) * g.dims[0] + cx;
                    for (int32_t k = g.cell_start[c]; k < g.cell_start[c + 1]; ++k) {
                        int32_t j = g.order[k];
                        rs.offer(dist2_f32(q, tgt + 4 * j), j);
                    }
                }
            }
        if (rs.full()) {
            // every point not yet scanned is farther than s*h along some axis (minus rounding slack)
            double covered = static_cast<double>(s) * g.h * (1.0 - 1e-5);
            if (static_cast<double>(rs.worst_d2()) < covered * covered) break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// probabilistic_weights.hpp
// ---------------------------------------------------------------------------------------------

struct WeightModel {
    // ProbabilisticWeights ctor, probabilistic_weights.hpp:30-46
    double v, t_exponent, log_norm_constant;
    int dimension;
    bool is_normal;
    WeightModel(double dof, int dim) : v(dof), t_exponent(0), log_norm_constant(0), dimension(dim), is_normal(false)
    {
        const double pi = std::atan(1.0) * 4;  // probabilistic_weights.hpp:13-16
        if (dof < std::numeric_limits<double>::infinity()) {
            t_exponent = -(dof + dim) / 2.0;
            // (v/2)*log(pi*v) is what the reference writes (hpp:41); it cancels in the row softmax.
            log_norm_constant = std::lgamma(dof / 2) - std::lgamma((dof + dim) / 2) + (dof / 2) * std::log(pi * dof);
        } else {
            is_normal = true;
            log_norm_constant = (dim / 2.0) * std::log(2 * pi);
        }
    }
    // updateWeights for one row, probabilistic_weights.hpp:56-101
    void row(const double* sq_err, int64_t k, double* out, std::vector<double>& log_probs,
             std::vector<double>& expected) const
    {
        log_probs.clear();
        expected.clear();
        double max_log_prob = -std::numeric_limits<double>::infinity();
        for (int64_t j = 0; j < k; ++j) {
            double lp;
            if (is_normal) {
                lp = -sq_err[j] / 2 + log_norm_constant;  // sign of the constant as in hpp:69
            } else {
                lp = t_exponent * std::log1p(sq_err[j] / v) - log_norm_constant;  // hpp:71-72
                expected.push_back((v + dimension) / (v + sq_err[j]));           // hpp:73
            }
            if (lp > max_log_prob) max_log_prob = lp;
            log_probs.push_back(lp);
        }
        double marginal = 0;
        for (double lp : log_probs) marginal += std::exp(lp - max_log_prob);  // hpp:83-85
        marginal = std::log(marginal) + max_log_prob;                         // hpp:86-87
        for (int64_t j = 0; j < k; ++j) {
            double w = std::exp(log_probs[j] - marginal);
            out[j] = is_normal ? w : w * expected[j];  // hpp:92-99
        }
    }
    void all_rows(int64_t n_rows, const int64_t* row_ptr, const double* sq_err, double* out) const
    {
        std::vector<double> lp, ex;
        for (int64_t i = 0; i < n_rows; ++i)
            row(sq_err + row_ptr[i], row_ptr[i + 1] - row_ptr[i], out + row_ptr[i], lp, ex);
    }
};

// ---------------------------------------------------------------------------------------------
// error_term.hpp residual, generic over the scalar so that dual numbers give AutoDiff's Jacobian
// ---------------------------------------------------------------------------------------------

struct Dual {  // value + 7 partials: [CERES] Jet<double,7>
    double a;
    double d[kNumParams];
};
inline Dual mk(double a)
{
    Dual r;
    r.a = a;
    for (int i = 0; i < kNumParams; ++i) r.d[i] = 0;
    return r;
}
inline Dual operator+(const Dual& x, const Dual& y)
{
    Dual r;
    r.a = x.a + y.a;
    for (int i = 0; i < kNumParams; ++i) r.d[i] = x.d[i] + y.d[i];
    return r;
}
inline Dual operator-(const Dual& x, const Dual& y)
{
    Dual r;
    r.a = x.a - y.a;
    for (int i = 0; i < kNumParams; ++i) r.d[i] = x.d[i] - y.d[i];
    return r;
}
inline Dual operator*(const Dual& x, const Dual& y)
{
    Dual r;
    r.a = x.a * y.a;
    for (int i = 0; i < kNumParams; ++i) r.d[i] = x.a * y.d[i] + x.d[i] * y.a;
    return r;
}
inline Dual operator/(const Dual& x, const Dual& y)
{
    Dual r;
    double inv = 1.0 / y.a;
    r.a = x.a * inv;
    for (int i = 0; i < kNumParams; ++i) r.d[i] = (x.d[i] - r.a * y.d[i]) * inv;
    return r;
}
inline Dual dsqrt(const Dual& x)
{
    Dual r;
    r.a = std::sqrt(x.a);
    double s = 0.5 / r.a;
    for (int i = 0; i < kNumParams; ++i) r.d[i] = x.d[i] * s;
    return r;
}
inline double dsqrt(double x) { return std::sqrt(x); }
inline double mk_like(const double&, double v) { return v; }
inline Dual mk_like(const Dual&, double v) { return mk(v); }

// [CERES] QuaternionRotatePoint: normalise q, then rotate with the unit-quaternion formula
//         p' = p + 2w (v x p) + 2 v x (v x p).
template <typename T>
inline void quaternion_rotate_point(const T q[4], const T pt[3], T out[3])
{
    T n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    T scale = mk_like(n2, 1.0) / dsqrt(n2);
    T u[4] = {q[0] * scale, q[1] * scale, q[2] * scale, q[3] * scale};
    T c0 = u[2] * pt[2] - u[3] * pt[1];
    T c1 = u[3] * pt[0] - u[1] * pt[2];
    T c2 = u[1] * pt[1] - u[2] * pt[0];
    c0 = c0 + c0;
    c1 = c1 + c1;
    c2 = c2 + c2;
    out[0] = pt[0] + u[0] * c0 + (u[2] * c2 - u[3] * c1);
    out[1] = pt[1] + u[0] * c1 + (u[3] * c0 - u[1] * c2);
    out[2] = pt[2] + u[0] * c2 + (u[1] * c1 - u[2] * c0);
}

// ErrorTerm::operator(), error_term.hpp:21-37: residual = y - (R(q) x + t); points promoted to double.
template <typename T>
inline void error_term(const float* src, const float* tgt, const T rot[4], const T trans[3], T res[3])
{
    T px[3] = {mk_like(rot[0], double(src[0])), mk_like(rot[0], double(src[1])), mk_like(rot[0], double(src[2]))};
    T py[3] = {mk_like(rot[0], double(tgt[0])), mk_like(rot[0], double(tgt[1])), mk_like(rot[0], double(tgt[2]))};
    T tp[3];
    quaternion_rotate_point(rot, px, tp);
    for (int i = 0; i < 3; ++i) {
        tp[i] = tp[i] + trans[i];
        res[i] = py[i] - tp[i];
    }
}

// ---------------------------------------------------------------------------------------------
// The per-outer-iteration problem (iteration.hpp:24-50) behind two interchangeable evaluators
// ---------------------------------------------------------------------------------------------

struct Association {
    const float* src;
    const float* tgt;
    int64_t n_rows;
    const int64_t* row_ptr;
    const int32_t* col;
    int64_t nnz() const { return row_ptr[n_rows]; }
};

struct ProblemBase {
    Association A;
    WeightModel wm;
    std::vector<double> weights;  // one ScaledLoss factor per residual block (error_term.hpp:39-43)
    std::vector<double> sq_err;
    int threads;
    ProblemBase(const Association& a, double dof, int th) : A(a), wm(dof, 3 /* DIMENSIONS, iteration.hpp:17 */), threads(th)
    {
        weights.assign(static_cast<size_t>(A.nnz()), 1.0);  // ScaledLoss(NULL, 1), error_term.hpp:17-19
        sq_err.resize(static_cast<size_t>(A.nnz()));
    }
    // WeightUpdaterCallback::operator(), weight_updater_callback.hpp:36-63
    void callback(const double x[kNumParams])
    {
        const int64_t n = A.n_rows;
#pragma omp parallel for schedule(static) num_threads(threads)
        for (int64_t i = 0; i < n; ++i) {
            for (int64_t k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
                double r[3];
                error_term<double>(A.src + 4 * i, A.tgt + 4 * static_cast<int64_t>(A.col[k]), x, x + 4, r);
                sq_err[k] = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            }
        }
#pragma omp parallel num_threads(threads)
        {
            std::vector<double> lp, ex;
#pragma omp for schedule(static)
            for (int64_t i = 0; i < n; ++i)
                wm.row(sq_err.data() + A.row_ptr[i], A.row_ptr[i + 1] - A.row_ptr[i], weights.data() + A.row_ptr[i], lp, ex);
        }
    }
    // [CERES] cost-only evaluation: 1/2 sum rho(|r|^2) with rho(s) = w s.
    double evaluate_cost(const double x[kNumParams]) const
    {
        const int64_t n = A.n_rows;
        double cost = 0;
#pragma omp parallel for schedule(static) reduction(+ : cost) num_threads(threads)
        for (int64_t i = 0; i < n; ++i) {
            for (int64_t k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
                double r[3];
                error_term<double>(A.src + 4 * i, A.tgt + 4 * static_cast<int64_t>(A.col[k]), x, x + 4, r);
                cost += 0.5 * weights[k] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            }
        }
        return cost;
    }
};

// Faithful evaluator: dual-number AutoDiff rows, dense 3K x 7 Jacobian, Householder QR on [J; D].
struct DenseQrProblem : ProblemBase {
    std::vector<double> J;  // row-major (3K) x 7, already scaled by sqrt(w) and the column scaling
    std::vector<double> f;  // 3K, scaled by sqrt(w)
    double grad[kNumParams];
    using ProblemBase::ProblemBase;

    double evaluate_full(const double x[kNumParams])
    {
        const int64_t K = A.nnz();
        J.assign(static_cast<size_t>(K) * 3 * kNumParams, 0.0);
        f.assign(static_cast<size_t>(K) * 3, 0.0);
        Dual X[kNumParams];
        for (int p = 0; p < kNumParams; ++p) {
            X[p] = mk(x[p]);
            X[p].d[p] = 1.0;
        }
        double cost = 0;
        for (int64_t i = 0; i < A.n_rows; ++i) {
            for (int64_t k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
                Dual r[3];
                error_term<Dual>(A.src + 4 * i, A.tgt + 4 * static_cast<int64_t>(A.col[k]), X, X + 4, r);
                double s = r[0].a * r[0].a + r[1].a * r[1].a + r[2].a * r[2].a;
                cost += 0.5 * weights[k] * s;
                // [CERES] Corrector with rho'' = 0: residual and Jacobian scaled by sqrt(rho') = sqrt(w)
                double sw = std::sqrt(weights[k]);
                for (int c = 0; c < 3; ++c) {
                    f[3 * k + c] = sw * r[c].a;
                    for (int p = 0; p < kNumParams; ++p) J[(3 * k + c) * kNumParams + p] = sw * r[c].d[p];
                }
            }
        }
        for (int p = 0; p < kNumParams; ++p) grad[p] = 0;
        for (int64_t row = 0; row < 3 * K; ++row)
            for (int p = 0; p < kNumParams; ++p) grad[p] += J[row * kNumParams + p] * f[row];
        return cost;
    }
    void squared_column_norms(double out[kNumParams]) const
    {
        for (int p = 0; p < kNumParams; ++p) out[p] = 0;
        const int64_t rows = static_cast<int64_t>(f.size());
        for (int64_t row = 0; row < rows; ++row)
            for (int p = 0; p < kNumParams; ++p) out[p] += J[row * kNumParams + p] * J[row * kNumParams + p];
    }
    void scale_columns(const double s[kNumParams])
    {
        const int64_t rows = static_cast<int64_t>(f.size());
        for (int64_t row = 0; row < rows; ++row)
            for (int p = 0; p < kNumParams; ++p) J[row * kNumParams + p] *= s[p];
    }
    void gradient(double g[kNumParams]) const
    {
        for (int p = 0; p < kNumParams; ++p) g[p] = grad[p];
    }
    // [CERES] DenseQRSolver: least squares on the Jacobian with diag(D) appended; returns y with J y ~ f.
    bool solve(const double D[kNumParams], double y[kNumParams]) const
    {
        const int64_t rows = static_cast<int64_t>(f.size());
        const int64_t R = rows + kNumParams;
        std::vector<double> M(static_cast<size_t>(R) * kNumParams);
        std::vector<double> b(static_cast<size_t>(R), 0.0);
        std::copy(J.begin(), J.end(), M.begin());
        std::copy(f.begin(), f.end(), b.begin());
        for (int p = 0; p < kNumParams; ++p)
            for (int c = 0; c < kNumParams; ++c) M[(rows + p) * kNumParams + c] = (p == c) ? D[p] : 0.0;
        // Householder QR, column by column, applied to b on the fly.
        for (int c = 0; c < kNumParams; ++c) {
            double norm2 = 0;
            for (int64_t r = c; r < R; ++r) norm2 += M[r * kNumParams + c] * M[r * kNumParams + c];
            double norm = std::sqrt(norm2);
            if (norm == 0.0) return false;
            double alpha = (M[c * kNumParams + c] > 0) ? -norm : norm;
            double v0 = M[c * kNumParams + c] - alpha;
            double vnorm2 = norm2 - M[c * kNumParams + c] * M[c * kNumParams + c] + v0 * v0;
            if (vnorm2 == 0.0) continue;
            // apply H = I - 2 v v^T / (v^T v) to the remaining columns and to b
            for (int cc = c + 1; cc < kNumParams; ++cc) {
                double dot = v0 * M[c * kNumParams + cc];
                for (int64_t r = c + 1; r < R; ++r) dot += M[r * kNumParams + c] * M[r * kNumParams + cc];
                double tau = 2.0 * dot / vnorm2;
                M[c * kNumParams + cc] -= tau * v0;
                for (int64_t r = c + 1; r < R; ++r) M[r * kNumParams + cc] -= tau * M[r * kNumParams + c];
            }
            double dotb = v0 * b[c];
            for (int64_t r = c + 1; r < R; ++r) dotb += M[r * kNumParams + c] * b[r];
            double taub = 2.0 * dotb / vnorm2;
            b[c] -= taub * v0;
            for (int64_t r = c + 1; r < R; ++r) b[r] -= taub * M[r * kNumParams + c];
            M[c * kNumParams + c] = alpha;
        }
        for (int c = kNumParams - 1; c >= 0; --c) {
            double s = b[c];
            for (int cc = c + 1; cc < kNumParams; ++cc) s -= M[c * kNumParams + cc] * y[cc];
            if (M[c * kNumParams + c] == 0.0) return false;
            y[c] = s / M[c * kNumParams + c];
        }
        return true;
    }
    // [CERES] model_cost_change = -(J step)^T (f + J step / 2)
    double model_cost_change(const double step[kNumParams]) const
    {
        const int64_t rows = static_cast<int64_t>(f.size());
        double acc = 0;
        for (int64_t row = 0; row < rows; ++row) {
            double m = 0;
            for (int p = 0; p < kNumParams; ++p) m += J[row * kNumParams + p] * step[p];
            acc += m * (f[row] + m / 2.0);
        }
        return -acc;
    }
};

// Fast evaluator: analytic Jacobian rows folded straight into the 7x7 normal equations (OpenMP).
// Mathematically the same LM step as DenseQrProblem; used for large clouds and as the CPU baseline.
struct NormalEqProblem : ProblemBase {
    double H[kNumParams][kNumParams];  // Js^T Js (column-scaled once scale_columns has been called)
    double g[kNumParams];              // Js^T f
    double grad_unscaled[kNumParams];
    using ProblemBase::ProblemBase;

    double evaluate_full(const double x[kNumParams])
    {
        // d(R(q/|q|) p)/dq = M(p) * (I - u u^T)/|q|,  M(p) = [2 b x p | -2a[p]x + 2((b.p)I + b p^T - 2 p b^T)]
        double n = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
        double u[4] = {x[0] / n, x[1] / n, x[2] / n, x[3] / n};
        double Pn[4][4];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) Pn[r][c] = ((r == c ? 1.0 : 0.0) - u[r] * u[c]) / n;
        const double a = u[0], b0 = u[1], b1 = u[2], b2 = u[3];
        const int64_t nrows = A.n_rows;
        const int T = threads;
        std::vector<double> part(static_cast<size_t>(T) * 64, 0.0);
#pragma omp parallel num_threads(T)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            double h[kNumParams][kNumParams] = {{0}};
            double gv[kNumParams] = {0};
            double cost = 0;
#pragma omp for schedule(static)
            for (int64_t i = 0; i < nrows; ++i) {
                if (A.row_ptr[i] == A.row_ptr[i + 1]) continue;
                const double p[3] = {double(A.src[4 * i]), double(A.src[4 * i + 1]), double(A.src[4 * i + 2])};
                double Mx[3][4];
                Mx[0][0] = 2 * (b1 * p[2] - b2 * p[1]);
                Mx[1][0] = 2 * (b2 * p[0] - b0 * p[2]);
                Mx[2][0] = 2 * (b0 * p[1] - b1 * p[0]);
                const double bp = b0 * p[0] + b1 * p[1] + b2 * p[2];
                const double bb[3] = {b0, b1, b2};
                const double skew[3][3] = {{0, -p[2], p[1]}, {p[2], 0, -p[0]}, {-p[1], p[0], 0}};
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        Mx[r][c + 1] = -2 * a * skew[r][c] + 2 * ((r == c ? bp : 0.0) + bb[r] * p[c] - 2 * p[r] * bb[c]);
                double Jrow[3][kNumParams];  // residual Jacobian = -[M Pn | I]
                for (int r = 0; r < 3; ++r) {
                    for (int c = 0; c < 4; ++c) {
                        double s = 0;
                        for (int k = 0; k < 4; ++k) s += Mx[r][k] * Pn[k][c];
                        Jrow[r][c] = -s;
                    }
                    for (int c = 0; c < 3; ++c) Jrow[r][4 + c] = (r == c) ? -1.0 : 0.0;
                }
                for (int64_t k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
                    double res[3];
                    error_term<double>(A.src + 4 * i, A.tgt + 4 * static_cast<int64_t>(A.col[k]), x, x + 4, res);
                    const double w = weights[k];
                    cost += 0.5 * w * (res[0] * res[0] + res[1] * res[1] + res[2] * res[2]);
                    for (int r = 0; r < 3; ++r) {
                        for (int c = 0; c < kNumParams; ++c) {
                            gv[c] += w * Jrow[r][c] * res[r];
                            for (int c2 = c; c2 < kNumParams; ++c2) h[c][c2] += w * Jrow[r][c] * Jrow[r][c2];
                        }
                    }
                }
            }
            double* out = part.data() + static_cast<size_t>(tid) * 64;
            int o = 0;
            for (int c = 0; c < kNumParams; ++c)
                for (int c2 = c; c2 < kNumParams; ++c2) out[o++] = h[c][c2];
            for (int c = 0; c < kNumParams; ++c) out[o++] = gv[c];
            out[o++] = cost;
        }
        double tot[64] = {0};
        for (int t = 0; t < T; ++t)
            for (int k = 0; k < 36; ++k) tot[k] += part[static_cast<size_t>(t) * 64 + k];
        int o = 0;
        for (int c = 0; c < kNumParams; ++c)
            for (int c2 = c; c2 < kNumParams; ++c2) {
                H[c][c2] = tot[o];
                H[c2][c] = tot[o++];
            }
        for (int c = 0; c < kNumParams; ++c) {
            g[c] = tot[o];
            grad_unscaled[c] = tot[o++];
        }
        return tot[o];
    }
    void squared_column_norms(double out[kNumParams]) const
    {
        for (int p = 0; p < kNumParams; ++p) out[p] = H[p][p];
    }
    void scale_columns(const double s[kNumParams])
    {
        for (int r = 0; r < kNumParams; ++r) {
            g[r] *= s[r];
            for (int c = 0; c < kNumParams; ++c) H[r][c] *= s[r] * s[c];
        }
    }
    void gradient(double out[kNumParams]) const
    {
        for (int p = 0; p < kNumParams; ++p) out[p] = grad_unscaled[p];
    }
    bool solve(const double D[kNumParams], double y[kNumParams]) const
    {
        double L[kNumParams][kNumParams];
        for (int r = 0; r < kNumParams; ++r)
            for (int c = 0; c < kNumParams; ++c) L[r][c] = H[r][c] + (r == c ? D[r] * D[r] : 0.0);
        for (int c = 0; c < kNumParams; ++c) {
            double d = L[c][c];
            for (int k = 0; k < c; ++k) d -= L[c][k] * L[c][k];
            if (!(d > 0.0)) return false;
            L[c][c] = std::sqrt(d);
            for (int r = c + 1; r < kNumParams; ++r) {
                double s = L[r][c];
                for (int k = 0; k < c; ++k) s -= L[r][k] * L[c][k];
                L[r][c] = s / L[c][c];
            }
        }
        double z[kNumParams];
        for (int r = 0; r < kNumParams; ++r) {
            double s = g[r];
            for (int k = 0; k < r; ++k) s -= L[r][k] * z[k];
            z[r] = s / L[r][r];
        }
        for (int r = kNumParams - 1; r >= 0; --r) {
            double s = z[r];
            for (int k = r + 1; k < kNumParams; ++k) s -= L[k][r] * y[k];
            y[r] = s / L[r][r];
        }
        return true;
    }
    double model_cost_change(const double step[kNumParams]) const
    {
        double lin = 0, quad = 0;
        for (int r = 0; r < kNumParams; ++r) {
            lin += step[r] * g[r];
            double s = 0;
            for (int c = 0; c < kNumParams; ++c) s += H[r][c] * step[c];
            quad += step[r] * s;
        }
        return -(lin + 0.5 * quad);
    }
};

// ---------------------------------------------------------------------------------------------
// [CERES] trust-region Levenberg-Marquardt, restated for exactly the options the reference sets
// (src/prob_point_cloud_registration.cc:88-98): DENSE_QR, use_nonmonotonic_steps, jacobi scaling,
// an IterationCallback that swaps the loss weights after every iteration (iteration.hpp:54-55).
// ---------------------------------------------------------------------------------------------

struct StepEvaluator {  // non-monotonic acceptance, Conn-Gould-Toint alg. 10.1.2, window 5
    double minimum_cost, current_cost, reference_cost, candidate_cost;
    double acc_reference_model_change = 0, acc_candidate_model_change = 0;
    int nonmonotonic = 0;
    int max_nonmonotonic;
    StepEvaluator(double c0, int maxn)
        : minimum_cost(c0), current_cost(c0), reference_cost(c0), candidate_cost(c0), max_nonmonotonic(maxn) {}
    double quality(double cost, double model_change) const
    {
        double rel = (current_cost - cost) / model_change;
        double hist = (reference_cost - cost) / (acc_reference_model_change + model_change);
        return std::max(rel, hist);
    }
    void accepted(double cost, double model_change)
    {
        current_cost = cost;
        acc_candidate_model_change += model_change;
        acc_reference_model_change += model_change;
        if (current_cost < minimum_cost) {
            minimum_cost = current_cost;
            nonmonotonic = 0;
            candidate_cost = current_cost;
            acc_candidate_model_change = 0;
        } else {
            ++nonmonotonic;
            if (current_cost > candidate_cost) {
                candidate_cost = current_cost;
                acc_candidate_model_change = 0;
            }
        }
        if (nonmonotonic == max_nonmonotonic) {
            reference_cost = candidate_cost;
            acc_reference_model_change = acc_candidate_model_change;
        }
    }
};

static std::atomic<long long> g_nonmonotonic_total{0};  // accepted non-monotonic steps since the last reset (test probe)

template <typename Problem>
void minimise(Problem& prob, double x[kNumParams], const oracle_solver_options& opt, oracle_solve_summary* sum)
{
    const double kInitialRadius = 1e4, kMaxRadius = 1e16, kMinRadius = 1e-32;
    const double kMinDiag = 1e-6, kMaxDiag = 1e32, kMinRelDecrease = 1e-3;
    const double kGradTol = 1e-10, kParamTol = 1e-8;
    const int kMaxInvalid = 5, kMaxNonmonotonic = 5;

    double scale[kNumParams], diag[kNumParams], D[kNumParams], step[kNumParams], grad[kNumParams];
    double best_x[kNumParams];
    const bool trace = std::getenv("PPCR_ORACLE_TRACE") != nullptr;

    // iteration zero
    double x_cost = prob.evaluate_full(x);
    prob.squared_column_norms(scale);
    for (int p = 0; p < kNumParams; ++p) scale[p] = 1.0 / (1.0 + std::sqrt(scale[p]));
    prob.scale_columns(scale);
    prob.gradient(grad);
    double grad_max = 0;
    for (int p = 0; p < kNumParams; ++p) grad_max = std::max(grad_max, std::fabs(grad[p]));
    double x_norm = 0;
    for (int p = 0; p < kNumParams; ++p) x_norm += x[p] * x[p];
    x_norm = std::sqrt(x_norm);

    sum->initial_cost = x_cost;
    double min_iteration_cost = x_cost;  // SetSummaryFinalCost: min over recorded iteration costs
    double minimum_cost = std::numeric_limits<double>::max();
    StepEvaluator ev(x_cost, kMaxNonmonotonic);
    double radius = kInitialRadius, decrease_factor = 2.0;
    bool reuse_diagonal = false, step_ok = true;
    int iteration = 0, invalid = 0, successful = 0, nonmonotonic_accepted = 0;
    int termination = 4;

    for (;;) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        if (step_ok) {
            ++successful;
            if (x_cost < minimum_cost) {
                minimum_cost = x_cost;
                std::copy(x, x + kNumParams, best_x);
            } else {
                ++nonmonotonic_accepted;  // an accepted step that did not lower the minimum: parameters_ stays behind x
                if (trace) std::fprintf(stderr, "  [oracle lm] it %d accepted non-monotonic step: x_cost %.12g >= minimum %.12g\n", iteration, x_cost, minimum_cost);
            }
        }
        // [CERES] update_state_every_iteration is a StateUpdatingCallback that runs ahead of the user callbacks and copies the
        // minimiser's `parameters_` into the user's arrays -- and `parameters_` is only overwritten when x_cost < minimum_cost
        // (the lines above).  After an accepted NON-monotonic step the WeightUpdaterCallback (weight_updater_callback.hpp:42-51
        // reads rotation_ / translation_) therefore sees the lowest-cost iterate, not x.
        prob.callback(best_x);
        if (iteration >= opt.max_num_iterations) { termination = 4; break; }
        if (step_ok && grad_max <= kGradTol) { termination = 2; break; }
        if (radius < kMinRadius) { termination = 3; break; }
        ++iteration;

        // ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep)
        if (!reuse_diagonal) {
            prob.squared_column_norms(diag);
            for (int p = 0; p < kNumParams; ++p) diag[p] = std::min(std::max(diag[p], kMinDiag), kMaxDiag);
        }
        for (int p = 0; p < kNumParams; ++p) D[p] = std::sqrt(diag[p] / radius);
        bool valid = prob.solve(D, step);
        reuse_diagonal = true;
        for (int p = 0; p < kNumParams && valid; ++p) valid = std::isfinite(step[p]);
        double model_change = 0;
        if (valid) {
            for (int p = 0; p < kNumParams; ++p) step[p] = -step[p];
            model_change = prob.model_cost_change(step);
            valid = model_change > 0.0;
        }
        if (!valid) {  // HandleInvalidStep
            if (++invalid >= kMaxInvalid) { termination = 6; break; }
            radius /= decrease_factor;
            decrease_factor *= 2.0;
            step_ok = false;
            min_iteration_cost = std::min(min_iteration_cost, x_cost);
            continue;
        }
        invalid = 0;
        double cand[kNumParams], step_norm = 0;
        for (int p = 0; p < kNumParams; ++p) {
            double delta = step[p] * scale[p];
            cand[p] = x[p] + delta;
            step_norm += (x[p] - cand[p]) * (x[p] - cand[p]);
        }
        step_norm = std::sqrt(step_norm);
        double cand_cost = prob.evaluate_cost(cand);

        if (trace && (step_norm <= kParamTol * (x_norm + kParamTol) || std::fabs(x_cost - cand_cost) <= opt.function_tolerance * x_cost))
            std::fprintf(stderr, "  [oracle lm] it %d terminating: x_cost %.12g cand %.12g step %.3g\n", iteration, x_cost, cand_cost, step_norm);
        if (step_norm <= kParamTol * (x_norm + kParamTol)) { termination = 1; break; }
        if (std::fabs(x_cost - cand_cost) <= opt.function_tolerance * x_cost) { termination = 0; break; }

        double quality = ev.quality(cand_cost, model_change);
        if (trace) std::fprintf(stderr, "  [oracle lm] it %d x_cost %.12g cand %.12g model %.6g quality %.6g radius %.4g step %.3g\n", iteration, x_cost, cand_cost, model_change, quality, radius, step_norm);
        if (quality > kMinRelDecrease) {  // HandleSuccessfulStep
            std::copy(cand, cand + kNumParams, x);
            x_norm = 0;
            for (int p = 0; p < kNumParams; ++p) x_norm += x[p] * x[p];
            x_norm = std::sqrt(x_norm);
            x_cost = prob.evaluate_full(x);
            prob.scale_columns(scale);
            prob.gradient(grad);
            grad_max = 0;
            for (int p = 0; p < kNumParams; ++p) grad_max = std::max(grad_max, std::fabs(grad[p]));
            step_ok = true;
            double t = 2.0 * quality - 1.0;
            radius = std::min(kMaxRadius, radius / std::max(1.0 / 3.0, 1.0 - t * t * t));
            decrease_factor = 2.0;
            reuse_diagonal = false;
            ev.accepted(cand_cost, model_change);
            min_iteration_cost = std::min(min_iteration_cost, x_cost);
        } else {
            step_ok = false;
            radius /= decrease_factor;
            decrease_factor *= 2.0;
            reuse_diagonal = true;
            min_iteration_cost = std::min(min_iteration_cost, cand_cost);
        }
    }
    std::copy(best_x, best_x + kNumParams, x);  // the minimiser hands back its lowest-cost iterate
    sum->final_cost = min_iteration_cost;
    sum->num_iterations = iteration;
    sum->num_successful_steps = successful;
    sum->termination = termination;
    sum->num_nonmonotonic_steps = nonmonotonic_accepted;
    g_nonmonotonic_total += nonmonotonic_accepted;
}

// iteration.hpp:59-67 + [EIGEN]: normalise q, rotation matrix, [R | t] as a row-major 4x4
void pose_to_matrix(const double q_in[4], const double t[3], double T[16])
{
    double n = std::sqrt(q_in[0] * q_in[0] + q_in[1] * q_in[1] + q_in[2] * q_in[2] + q_in[3] * q_in[3]);
    double w = q_in[0] / n, x = q_in[1] / n, y = q_in[2] / n, z = q_in[3] / n;
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w;
    double txx = tx * x, txy = ty * x, txz = tz * x;
    double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[4 * r + c] = R[3 * r + c];
        T[4 * r + 3] = t[r];
    }
    T[12] = T[13] = T[14] = 0;
    T[15] = 1;
}

int solve_iteration(const Association& A, const oracle_params& P, const oracle_solver_options& O, double rot[4],
                    double trans[3], double T[16], oracle_solve_summary* sum)
{
    double x[kNumParams];
    std::copy(P.initial_rotation, P.initial_rotation + 4, x);       // iteration.hpp:31-34
    std::copy(P.initial_translation, P.initial_translation + 3, x + 4);
    std::memset(sum, 0, sizeof(*sum));
    if (A.nnz() == 0) {
        // [CERES] a problem without residual blocks returns at once with zero costs
        sum->termination = 5;
    } else if (O.inner_kind == 0) {
        DenseQrProblem prob(A, P.dof, 1);
        prob.callback(x);  // iteration.hpp:49
        minimise(prob, x, O, sum);
    } else {
        NormalEqProblem prob(A, P.dof, resolve_threads(O.num_threads));
        prob.callback(x);
        minimise(prob, x, O, sum);
    }
    std::copy(x, x + 4, rot);
    std::copy(x + 4, x + 7, trans);
    pose_to_matrix(rot, trans, T);
    return 0;
}

// [PCL] transformPointCloud(Affine3d): double arithmetic, row by row, rounded to float on store.
void transform_cloud(float* pts, int64_t n, const double* T)
{
    for (int64_t i = 0; i < n; ++i) {
        double x = pts[4 * i], y = pts[4 * i + 1], z = pts[4 * i + 2];
        for (int r = 0; r < 3; ++r) pts[4 * i + r] = static_cast<float>(T[4 * r] * x + T[4 * r + 1] * y + T[4 * r + 2] * z + T[4 * r + 3]);
    }
}

void matmul4(const double* A, const double* B, double* C)
{
    double tmp[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * r + k] * B[4 * k + c];
            tmp[4 * r + c] = s;
        }
    std::copy(tmp, tmp + 16, C);
}

// [PCL] VoxelGrid<PointXYZ>::applyFilter with default settings (no field filter, min 0 points/voxel).
int64_t voxel_grid(const float* in, int64_t n, double leaf_d, float* out)
{
    if (n == 0) return 0;
    const float leaf = static_cast<float>(leaf_d);  // setLeafSize(float, float, float)
    const float inv = 1.0f / leaf;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < n; ++i) {
        if (!std::isfinite(in[4 * i]) || !std::isfinite(in[4 * i + 1]) || !std::isfinite(in[4 * i + 2])) continue;
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], in[4 * i + a]);
            hi[a] = std::max(hi[a], in[4 * i + a]);
        }
    }
    int64_t d[3];
    int32_t minb[3];
    for (int a = 0; a < 3; ++a) {
        d[a] = static_cast<int64_t>((hi[a] - lo[a]) * inv) + 1;
        minb[a] = static_cast<int32_t>(std::floor(lo[a] * inv));
        int32_t maxb = static_cast<int32_t>(std::floor(hi[a] * inv));
        (void)maxb;
    }
    // PCL multiplies the three int64 extents; the product is formed in double here so that absurd extents (where
    // PCL's own multiplication would wrap) still take the "too small" branch.  Exact wherever the test is close.
    if (static_cast<double>(d[0]) * static_cast<double>(d[1]) * static_cast<double>(d[2]) >
        static_cast<double>(std::numeric_limits<int32_t>::max())) {
        std::memcpy(out, in, static_cast<size_t>(n) * 16);  // "Leaf size is too small": output = input
        return -1;
    }
    int32_t div[3];
    for (int a = 0; a < 3; ++a) {
        int32_t maxb = static_cast<int32_t>(std::floor(hi[a] * inv));
        div[a] = maxb - minb[a] + 1;
    }
    const int32_t mul[3] = {1, div[0], div[0] * div[1]};
    struct Entry {
        uint32_t voxel;
        uint32_t point;
    };
    std::vector<Entry> entries;
    entries.reserve(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
        if (!std::isfinite(in[4 * i]) || !std::isfinite(in[4 * i + 1]) || !std::isfinite(in[4 * i + 2])) continue;
        int32_t ijk[3];
        for (int a = 0; a < 3; ++a) ijk[a] = static_cast<int32_t>(std::floor(in[4 * i + a] * inv) - static_cast<float>(minb[a]));
        uint32_t v = static_cast<uint32_t>(ijk[0] * mul[0] + ijk[1] * mul[1] + ijk[2] * mul[2]);
        entries.push_back({v, static_cast<uint32_t>(i)});
    }
    // PCL sorts on the voxel id only; a stable sort fixes the (otherwise unspecified) in-voxel order.
    std::stable_sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& b) { return a.voxel < b.voxel; });
    int64_t n_out = 0;
    size_t k = 0;
    while (k < entries.size()) {
        size_t e = k;
        float cx = 0, cy = 0, cz = 0;  // float32 centroid accumulation, like PCL's Eigen::Vector4f
        while (e < entries.size() && entries[e].voxel == entries[k].voxel) {
            const float* p = in + 4 * static_cast<size_t>(entries[e].point);
            cx += p[0];
            cy += p[1];
            cz += p[2];
            ++e;
        }
        float cnt = static_cast<float>(e - k);
        out[4 * n_out] = cx / cnt;
        out[4 * n_out + 1] = cy / cnt;
        out[4 * n_out + 2] = cz / cnt;
        out[4 * n_out + 3] = 1.0f;
        ++n_out;
        k = e;
    }
    return n_out;
}

int64_t radius_search(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, double radius, int32_t max_nn,
                      int32_t cap, int use_grid, int threads, int32_t* out_idx, float* out_d2, int32_t* out_count)
{
    const float r2f = static_cast<float>(radius * radius);  // pcl::KdTreeFLANN::radiusSearch passes radius*radius
    const size_t limit = effective_limit(max_nn, n_tgt);
    CpuGrid grid;
    if (use_grid && n_tgt > 0) build_grid(tgt, n_tgt, radius, limit, grid);
    int64_t total = 0;
    const int T = resolve_threads(threads);
#pragma omp parallel num_threads(T) reduction(+ : total)
    {
        ResultSet rs;
        std::vector<DistIndex> sorted;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n_src; ++i) {
            rs.reset(limit, r2f);
            const float* q = src + 4 * i;
            if (use_grid && n_tgt > 0) {
                grid_query(grid, tgt, q, r2f, radius, rs);
            } else {
                for (int64_t j = 0; j < n_tgt; ++j) rs.offer(dist2_f32(q, tgt + 4 * j), static_cast<int32_t>(j));
            }
            rs.sorted(sorted);
            size_t k = std::min(sorted.size(), static_cast<size_t>(cap));
            out_count[i] = static_cast<int32_t>(k);
            for (size_t c = 0; c < k; ++c) {
                out_idx[i * cap + static_cast<int64_t>(c)] = sorted[c].idx;
                if (out_d2) out_d2[i * cap + static_cast<int64_t>(c)] = sorted[c].d2;
            }
            total += static_cast<int64_t>(k);
        }
    }
    return total;
}


// ---------------------------------------------------------------------------------------------
// utilities.hpp:28-234 -- the closest-point metric helpers (1-nearest-neighbour distances of cloud1 in cloud2)
// ---------------------------------------------------------------------------------------------

// [PCL/FLANN] KdTreeFLANN::nearestKSearch(cloud, i, 1, idx, dist): dist[0] is the SQUARED distance (float, L2_Simple
// arithmetic) to the nearest point of the tree's cloud.  Exact search: brute force returns the same value.
static void closest_sq_distances(const float* c1, int64_t n1, const float* c2, int64_t n2, std::vector<float>& out)
{
    out.assign(static_cast<size_t>(n1), 0.f);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n1; ++i) {
        float best = std::numeric_limits<float>::infinity();
        for (int64_t j = 0; j < n2; ++j) best = std::min(best, dist2_f32(c1 + 4 * i, c2 + 4 * j));
        out[static_cast<size_t>(i)] = best;
    }
}

// The reference's "median" of a SORTED vector (utilities.hpp:84-89 and five more copies): element (n + 1) / 2 when n is odd,
// the mean of elements n / 2 and n / 2 + 1 when n is even -- one position above the textbook median, and out of bounds for
// n = 1 and n = 2 (undefined behaviour in the reference; NaN here).  T = double for the vector<double> helpers (the addition
// happens in double), float for the vector<float> ones (the two floats are added in float, then divided by 2.0 in double).
template <typename T>
static double reference_median(const std::vector<T>& v)
{
    const size_t n = v.size();
    if (n % 2 != 0) {
        const size_t k = (n + 1) / 2;
        return k < n ? static_cast<double>(v[k]) : std::numeric_limits<double>::quiet_NaN();
    }
    const size_t a = n / 2, b = n / 2 + 1;
    if (b >= n) return std::numeric_limits<double>::quiet_NaN();
    return (v[a] + v[b]) / 2.0;
}


}  // namespace

extern "C" {

int32_t oracle_closest_metrics(const float* c1, int64_t n1, const float* c2, int64_t n2, double factor, double* out,
                                          float* out_d2)
{
    if (n1 < 1 || n2 < 1) return -1;
    std::vector<float> d;
    closest_sq_distances(c1, n1, c2, n2, d);
    if (out_d2) std::copy(d.begin(), d.end(), out_d2);
    // averageClosestDistance :28-46 and sumSquaredError :48-65: float distances added to a double in point order
    double sum = 0;
    for (float v : d) sum += v;
    out[0] = sum / static_cast<double>(n1);
    out[1] = sum;
    // robustSumSquaredError :67-101 / (factor) :103-138 / robustAveragedSumSquaredError :140-175: vector<double>, sorted
    std::vector<double> all(d.begin(), d.end());
    std::sort(all.begin(), all.end());
    const double med = reference_median(all);
    auto robust = [&](double f, double* s_out, int64_t* n_out) {
        double s = 0;
        int64_t nf = 0;
        for (double v : all)
            if (v <= med * f && v >= med / f) {
                s += v;
                ++nf;
            }
        *s_out = s;
        *n_out = nf;
    };
    double s3, sf;
    int64_t n3, nf;
    robust(3.0, &s3, &n3);
    robust(factor, &sf, &nf);
    const double big = std::numeric_limits<double>::max();
    out[2] = n3 < 10 ? big : s3;
    out[3] = nf < 10 ? big : sf;
    out[4] = n3 < 10 ? big : s3 / static_cast<double>(n3);
    // medianClosestDistance :177-199 and robustMedianClosestDistance :201-234: vector<float>, sorted
    std::vector<float> fs(d);
    std::sort(fs.begin(), fs.end());
    const double medf = reference_median(fs);
    out[5] = medf;
    std::vector<float> filt;
    for (float v : fs)
        if (v <= medf * 3 && v >= medf / 3.0) filt.push_back(v);
    out[6] = filt.empty() ? std::numeric_limits<double>::quiet_NaN() : reference_median(filt) / static_cast<double>(filt.size());
    out[7] = static_cast<double>(n3);
    out[8] = static_cast<double>(nf);
    return 0;
}

// H (upper triangle, 28), g (7) and the cost of one Jacobian evaluation at pose x_e with the weights refreshed at pose x_w:
// what Ceres assembles from the 3K x 7 Jacobian of the problem of iteration.hpp:24-50 (unscaled columns).
void oracle_normal_eq(const float* src, const float* tgt, int64_t n_rows, const int64_t* row_ptr, const int32_t* col_idx,
                                 double dof, const double* x_w, const double* x_e, double* out36)
{
    Association A{src, tgt, n_rows, row_ptr, col_idx};
    NormalEqProblem prob(A, dof, 1);
    prob.callback(x_w);
    const double cost = prob.evaluate_full(x_e);
    int o = 0;
    for (int r = 0; r < kNumParams; ++r)
        for (int c = r; c < kNumParams; ++c) out36[o++] = prob.H[r][c];
    for (int r = 0; r < kNumParams; ++r) out36[o++] = prob.g[r];
    out36[o] = cost;
}


int32_t oracle_max_threads(void) { return resolve_threads(0); }

int64_t oracle_nonmonotonic_steps(int32_t reset)
{
    const long long v = g_nonmonotonic_total.load();
    if (reset) g_nonmonotonic_total.store(0);
    return v;
}

int64_t oracle_radius_search(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, double radius,
                             int32_t max_nn, int32_t cap, int32_t use_grid, int32_t num_threads, int32_t* out_idx,
                             float* out_d2, int32_t* out_count)
{
    return radius_search(src, n_src, tgt, n_tgt, radius, max_nn, cap, use_grid, num_threads, out_idx, out_d2, out_count);
}

void oracle_update_weights(int64_t n_rows, const int64_t* row_ptr, const double* squared_errors, double dof,
                           int32_t dimension, double* out_weights)
{
    WeightModel wm(dof, dimension);
    wm.all_rows(n_rows, row_ptr, squared_errors, out_weights);
}

void oracle_callback_weights(const float* src, const float* tgt, int64_t n_rows, const int64_t* row_ptr,
                             const int32_t* col_idx, const double* rotation, const double* translation, double dof,
                             double* out_sq_err, double* out_weights)
{
    Association A{src, tgt, n_rows, row_ptr, col_idx};
    ProblemBase pb(A, dof, 1);
    double x[kNumParams];
    std::copy(rotation, rotation + 4, x);
    std::copy(translation, translation + 3, x + 4);
    pb.callback(x);
    if (out_sq_err) std::copy(pb.sq_err.begin(), pb.sq_err.end(), out_sq_err);
    std::copy(pb.weights.begin(), pb.weights.end(), out_weights);
}

int32_t oracle_iteration_solve(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, const int64_t* row_ptr,
                               const int32_t* col_idx, const oracle_params* params, const oracle_solver_options* opts,
                               double* out_rotation, double* out_translation, double* out_T, oracle_solve_summary* summary)
{
    (void)n_tgt;
    Association A{src, tgt, n_src, row_ptr, col_idx};
    return solve_iteration(A, *params, *opts, out_rotation, out_translation, out_T, summary);
}

void oracle_transform(float* xyzw, int64_t n, const double* T) { transform_cloud(xyzw, n, T); }

int64_t oracle_voxel_grid(const float* xyzw, int64_t n, double leaf, float* out) { return voxel_grid(xyzw, n, leaf, out); }

double oracle_calculate_mse(const float* a, const float* b, int64_t n)
{
    // utilities.hpp:16-26; pcl::euclideanDistance works in float
    double acc = 0;
    for (int64_t i = 0; i < n; ++i) {
        float dx = a[4 * i] - b[4 * i], dy = a[4 * i + 1] - b[4 * i + 1], dz = a[4 * i + 2] - b[4 * i + 2];
        acc += std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    return acc / static_cast<double>(n);
}

int32_t oracle_align(const float* src_in, int64_t n_src_in, const float* tgt_in, int64_t n_tgt_in,
                     const oracle_params* params, const oracle_solver_options* opts, int32_t use_grid, double* history,
                     oracle_iter_stats* stats, int32_t max_hist, float* out_filtered_source, int64_t* n_filtered_src,
                     int64_t* n_filtered_tgt)
{
    const oracle_params& P = *params;
    // ctor, src/prob_point_cloud_registration.cc:15-49
    std::vector<float> source(src_in, src_in + 4 * n_src_in);
    std::vector<float> target(tgt_in, tgt_in + 4 * n_tgt_in);
    int64_t n_src = n_src_in, n_tgt = n_tgt_in;
    if (P.source_filter_size > 0) {
        std::vector<float> tmp(static_cast<size_t>(4 * n_src));
        int64_t k = voxel_grid(source.data(), n_src, P.source_filter_size, tmp.data());
        if (k >= 0) {
            n_src = k;
            tmp.resize(static_cast<size_t>(4 * k));
            source.swap(tmp);
        }
    }
    if (P.target_filter_size > 0) {
        std::vector<float> tmp(static_cast<size_t>(4 * n_tgt));
        int64_t k = voxel_grid(target.data(), n_tgt, P.target_filter_size, tmp.data());
        if (k >= 0) {
            n_tgt = k;
            tmp.resize(static_cast<size_t>(4 * k));
            target.swap(tmp);
        }
    }
    if (n_filtered_src) *n_filtered_src = n_src;
    if (n_filtered_tgt) *n_filtered_tgt = n_tgt;

    // align(), src/prob_point_cloud_registration.cc:63-136 with hasConverged() :138-158
    int current_iteration = 0;
    double cost_drop = 0;
    int num_unuseful = 0;
    double prev_T[16];
    const size_t limit = effective_limit(P.max_neighbours, n_tgt);
    const int32_t cap = static_cast<int32_t>(std::min<size_t>(limit, static_cast<size_t>(std::max<int64_t>(n_tgt, 1))));
    std::vector<int32_t> idx(static_cast<size_t>(n_src) * cap), count(static_cast<size_t>(n_src));
    std::vector<int64_t> row_ptr(static_cast<size_t>(n_src) + 1);
    std::vector<int32_t> col;
    for (;;) {
        // hasConverged
        if (current_iteration == P.n_iter) break;
        if (cost_drop < P.cost_drop_thresh) {
            if (num_unuseful > P.n_cost_drop_it) break;
            ++num_unuseful;
        } else {
            num_unuseful = 0;
        }
        // data association, :66-83.  Eigen's setFromTriplets leaves each row sorted by column.
        radius_search(source.data(), n_src, target.data(), n_tgt, P.radius, P.max_neighbours, cap, use_grid,
                      opts->num_threads, idx.data(), nullptr, count.data());
        row_ptr[0] = 0;
        for (int64_t i = 0; i < n_src; ++i) row_ptr[i + 1] = row_ptr[i] + count[i];
        col.resize(static_cast<size_t>(row_ptr[n_src]));
        for (int64_t i = 0; i < n_src; ++i) {
            std::copy(idx.begin() + i * cap, idx.begin() + i * cap + count[i], col.begin() + row_ptr[i]);
            std::sort(col.begin() + row_ptr[i], col.begin() + row_ptr[i + 1]);
        }
        Association A{source.data(), target.data(), n_src, row_ptr.data(), col.data()};
        double rot[4], trans[3], dT[16], cur[16];
        oracle_solve_summary sum;
        solve_iteration(A, P, *opts, rot, trans, dT, &sum);  // :85-100
        if (current_iteration > 0) matmul4(dT, prev_T, cur); else std::copy(dT, dT + 16, cur);  // :101-107
        std::copy(cur, cur + 16, prev_T);
        if (current_iteration < max_hist) {
            if (history) std::copy(cur, cur + 16, history + 16 * current_iteration);
            if (stats) {
                stats[current_iteration].initial_cost = sum.initial_cost;
                stats[current_iteration].final_cost = sum.final_cost;
                stats[current_iteration].n_correspondences = row_ptr[n_src];
                stats[current_iteration].lm_iterations = sum.num_iterations;
                stats[current_iteration].num_successful_steps = sum.num_successful_steps;
            }
        }
        transform_cloud(source.data(), n_src, dT);                               // :110-112
        cost_drop = (sum.initial_cost - sum.final_cost) / sum.initial_cost;      // :119 (NaN when 0/0)
        if (stats && current_iteration < max_hist) stats[current_iteration].cost_drop = cost_drop;
        ++current_iteration;                                                    // :130
    }
    if (out_filtered_source) std::copy(source.begin(), source.end(), out_filtered_source);
    return current_iteration;
}

}  // extern "C"
