// ppcr_lm.h -- the scalar part of one registration: moment expansion into the 7x7 normal equations, the
// Levenberg-Marquardt controller, pose composition and the outer convergence test.
//
// Everything here is plain C++ qualified PPCR_HD so that the single-CTA controller kernel (ppcr_kernels.cu) and
// the CPU unit tests of the host logic (tests/emu) run the same source (one device-only shortcut: inv_sqrt).
//
// What it replaces in the reference (paths relative to the reference tree):
//   * ceres::Solve on the problem built at prob_point_cloud_registration_iteration.hpp:24-57 with the options
//     of src/prob_point_cloud_registration.cc:88-98 (trust-region LM, DENSE_QR, non-monotonic steps,
//     Jacobi scaling, an IterationCallback that refreshes the loss weights after every iteration);
//   * ProbPointCloudRegistrationIteration::transformation(), iteration.hpp:59-67;
//   * the pose composition, cost-drop and hasConverged() of src/prob_point_cloud_registration.cc:101-158.
//
// Reformulation.  The residual of correspondence (i,j) is r_ij = y_j - (R(q/|q|) x_i + t) (error_term.hpp:21-37);
// its Jacobian -[M(x_i) Pn | I] depends only on the source point, M(x) is linear in x and Pn = (I - u u^T)/|q|.
// So J^T W J, J^T W r and the cost follow from 23 weighted moments of the source points,
//     S0 = sum W_i, S1 = sum W_i x_i, S2 = sum W_i x_i x_i^T, Sr = sum rho_i, C = sum x_i rho_i^T, cost,
// with W_i = sum_j w_ij, rho_i = sum_j w_ij r_ij.  The eval kernel streams the association once and produces the
// moments; this file turns them into the dense 7x7 system Ceres would have built from its 3K x 7 Jacobian.
#ifndef PPCR_LM_H
#define PPCR_LM_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PPCR_HD __host__ __device__ inline
#else
#define PPCR_HD inline
#endif

namespace ppcr {

constexpr int kNP = 7;     // rotation_[4] (w,x,y,z) + translation_[3]
constexpr int kNSum = 24;  // moments produced by the eval kernel (one spare slot keeps rows 64-bit x 24)

// layout of the moment vector
enum MomentSlot {
    M_S0 = 0,     // sum W_i
    M_S1 = 1,     // 1..3   sum W_i x_i
    M_S2 = 4,     // 4..9   sum W_i x x^T : xx xy xz yy yz zz
    M_SR = 10,    // 10..12 sum rho_i
    M_C = 13,     // 13..21 sum x_ic rho_id, row-major [c][d]
    M_COST = 22,  // 1/2 sum_ij w_ij |r_ij|^2
    M_ROWS = 23   // number of source rows with at least one neighbour (diagnostic)
};

enum Phase { PH_SEARCH = 0, PH_LM = 1, PH_DONE = 2 };

enum Termination {
    TERM_FUNCTION_TOL = 0, TERM_PARAMETER_TOL = 1, TERM_GRADIENT_TOL = 2, TERM_MIN_RADIUS = 3,
    TERM_MAX_ITER = 4, TERM_NO_RESIDUALS = 5, TERM_INVALID_STEPS = 6
};

struct Pose {  // R(q/|q|) row-major and t: what the eval kernel needs to move a source point
    double R[9];
    double t[3];
};

struct Config {  // immutable per registration
    double x0[kNP];              // params.initial_rotation / initial_translation
    double function_tolerance;   // 10e-6 in align() (registration.cc:97)
    double cost_drop_thresh;
    double n_cost_drop_it;
    double dof;
    int32_t n_iter;
    int32_t max_lm_iterations;   // INT_MAX in the reference (registration.cc:96)
    int32_t is_normal;           // dof == +inf
    int32_t fast_weights;
};

struct IterStats {  // same layout as ppcr_iter_stats
    double initial_cost, final_cost, cost_drop;
    int64_t n_correspondences;
    int32_t lm_iterations, num_successful_steps;
};

struct PairState {
    // ---- inner LM (one ceres::Solve) ----
    double x[kNP], cand[kNP], best_x[kNP];
    Pose pose_e;  // pose the NEXT eval computes residuals at (the candidate)
    Pose pose_w;  // pose the NEXT eval refreshes the weights at (the lowest-cost iterate so far: what Ceres shows the callback)
    double Hs[kNP * kNP], gs[kNP];  // column-scaled J^T W J and J^T W r of the current iterate
    double scale[kNP], diag[kNP], step[kNP];
    double x_cost, x_norm, grad_max, minimum_cost, min_iter_cost, initial_cost;
    double radius, decrease_factor, model_change;
    double ev_min, ev_cur, ev_ref, ev_cand, ev_acc_ref, ev_acc_cand;  // non-monotonic step evaluator
    int32_t ev_nonmono;
    int32_t iteration, invalid, successful, reuse_diag, step_ok, termination;
    // ---- outer loop ----
    double T_total[16], dT[16];
    double cost_drop;
    int32_t current_iteration, num_unuseful, phase, apply_dT;
    int64_t K;        // correspondences of the current association (search kernel)
    int64_t K_total;  // summed over outer iterations
    int64_t exchange_cycles;  // sharded pairs: SM clocks the controller block spent in the moment exchange (send + wait), summed
    int32_t ticks, evals;
    int32_t error;
    int32_t search_cursor;  // next chunk of queries to hand out (persistent search kernel)
    int32_t eval_ticket;    // blocks of the eval kernel that have published their partial sums
    int32_t row_overflow;   // set by a search kernel: a row reached PairDev::overflow_at
};

// ------------------------------------------------------------------------------------------------------------

PPCR_HD void pose_from_x(const double* x, Pose* p)
{
    // one reciprocal instead of four divisions: this runs on a single device thread between two passes over the cloud
    const double inv_n = 1.0 / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
    const double a = x[0] * inv_n, b0 = x[1] * inv_n, b1 = x[2] * inv_n, b2 = x[3] * inv_n;
    // R = I + 2a[b]x + 2[b]x^2  (the unit-quaternion rotation Ceres applies after normalising)
    p->R[0] = 1.0 - 2.0 * (b1 * b1 + b2 * b2);
    p->R[1] = 2.0 * (b0 * b1 - a * b2);
    p->R[2] = 2.0 * (b0 * b2 + a * b1);
    p->R[3] = 2.0 * (b0 * b1 + a * b2);
    p->R[4] = 1.0 - 2.0 * (b0 * b0 + b2 * b2);
    p->R[5] = 2.0 * (b1 * b2 - a * b0);
    p->R[6] = 2.0 * (b0 * b2 - a * b1);
    p->R[7] = 2.0 * (b1 * b2 + a * b0);
    p->R[8] = 1.0 - 2.0 * (b0 * b0 + b1 * b1);
    p->t[0] = x[4];
    p->t[1] = x[5];
    p->t[2] = x[6];
}

// iteration.hpp:59-67 with Eigen's normalize()/toRotationMatrix(): [R | t] as a row-major 4x4
PPCR_HD void matrix_from_x(const double* x, double* T)
{
    const double inv_n = 1.0 / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
    const double w = x[0] * inv_n, qx = x[1] * inv_n, qy = x[2] * inv_n, qz = x[3] * inv_n;
    const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    T[0] = 1.0 - (tyy + tzz); T[1] = txy - twz;         T[2] = txz + twy;          T[3] = x[4];
    T[4] = txy + twz;         T[5] = 1.0 - (txx + tzz); T[6] = tyz - twx;          T[7] = x[5];
    T[8] = txz - twy;         T[9] = tyz + twx;         T[10] = 1.0 - (txx + tyy); T[11] = x[6];
    T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;
}

PPCR_HD void matmul4(const double* A, const double* B, double* C)
{
    double tmp[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[4 * r + k] * B[4 * k + c];
            tmp[4 * r + c] = s;
        }
    for (int k = 0; k < 16; ++k) C[k] = tmp[k];
}

// ---- moments -> dense J^T W J (H, 7x7 row-major), J^T W r (g) and cost at the pose x the residuals were taken at ----
//
// Split into independent pieces so that the controller block can spread them over its threads (36 entries of N, then
// 27 output tasks) while the CPU build runs the same pieces in a loop: both produce the same bits.

struct Expanded {
    double H[kNP * kNP];
    double g[kNP];
    double cost;
};

struct QuatFrame {  // u = q / |q| and 1 / |q|
    double u[4];
    double inv_n;
};

PPCR_HD void quat_frame(const double* x, QuatFrame* f)
{
    const double n = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
    f->inv_n = 1.0 / n;
    for (int k = 0; k < 4; ++k) f->u[k] = x[k] * f->inv_n;
}

constexpr int kExpandN = 36;      // entries of N[c][r][k], index (c * 3 + r) * 4 + k
constexpr int kExpandTasks = 27;  // 10 H_qq pairs, 12 H_qt entries, 4 g_q entries, 1 task for H_tt / g_t / cost

// N_c = M(e_c) * Pn with M(p) = [ 2 b x p | -2a[p]x + 2((b.p) I + b p^T - 2 p b^T) ], Pn = (I - u u^T) / |q|
PPCR_HD double expand_N_entry(const QuatFrame& f, int idx)
{
    const int c = idx / 12, r = (idx / 4) % 3, k = idx % 4;
    const double a = f.u[0];
    const double b[3] = {f.u[1], f.u[2], f.u[3]};
    double e[3] = {0.0, 0.0, 0.0};
    e[c] = 1.0;
    double M[4];
    const int r1 = (r + 1) % 3, r2 = (r + 2) % 3;
    M[0] = 2.0 * (b[r1] * e[r2] - b[r2] * e[r1]);  // 2 (b x e)_r
    for (int kk = 0; kk < 3; ++kk) {
        // skew(e)[r][kk]: +e[j] / -e[j] on the off-diagonals
        double sk = 0.0;
        if (kk == r1) sk = -e[r2];
        else if (kk == r2) sk = e[r1];
        M[kk + 1] = -2.0 * a * sk + 2.0 * ((r == kk ? b[c] : 0.0) + b[r] * e[kk] - 2.0 * e[r] * b[kk]);
    }
    double s = 0.0;
    for (int l = 0; l < 4; ++l) s += M[l] * (((l == k ? 1.0 : 0.0) - f.u[l] * f.u[k]) * f.inv_n);
    return s;
}

PPCR_HD void expand_task(const double* S, const double* N, int task, Expanded* out)
{
    if (task < 10) {  // H_qq(p, q) = sum_{c,d} S2[c][d] N_c[:,p] . N_d[:,q]
        int p = 0, q = task;
        while (q >= 4 - p) {
            q -= 4 - p;
            ++p;
        }
        q += p;
        const int s2i[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        double s = 0.0;
        for (int c = 0; c < 3; ++c)
            for (int d = 0; d < 3; ++d) {
                double nn = 0.0;
                for (int r = 0; r < 3; ++r) nn += N[(c * 3 + r) * 4 + p] * N[(d * 3 + r) * 4 + q];
                s += S[M_S2 + s2i[c][d]] * nn;
            }
        out->H[p * kNP + q] = s;
        out->H[q * kNP + p] = s;
    } else if (task < 22) {  // H_qt(p, r) = sum_c S1[c] N_c[r][p]
        const int p = (task - 10) / 3, r = (task - 10) % 3;
        double s = 0.0;
        for (int c = 0; c < 3; ++c) s += S[M_S1 + c] * N[(c * 3 + r) * 4 + p];
        out->H[p * kNP + 4 + r] = s;
        out->H[(4 + r) * kNP + p] = s;
    } else if (task < 26) {  // g_q(p) = -sum_{c,r} N_c[r][p] C[c][r]
        const int p = task - 22;
        double s = 0.0;
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) s += N[(c * 3 + r) * 4 + p] * S[M_C + 3 * c + r];
        out->g[p] = -s;
    } else {  // H_tt = S0 I, g_t = -S_rho, cost
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) out->H[(4 + r) * kNP + 4 + c] = (r == c) ? S[M_S0] : 0.0;
        for (int r = 0; r < 3; ++r) out->g[4 + r] = -S[M_SR + r];
        out->cost = S[M_COST];
    }
}

PPCR_HD void expand_moments(const double* S, const double* x, Expanded* out)
{
    QuatFrame f;
    quat_frame(x, &f);
    double N[kExpandN];
    for (int k = 0; k < kExpandN; ++k) N[k] = expand_N_entry(f, k);
    for (int t = 0; t < kExpandTasks; ++t) expand_task(S, N, t, out);
}

// 1 / sqrt(d): one reciprocal square root on the device, sqrt + division on the host (last-bit differences only)
PPCR_HD double inv_sqrt(double d)
{
#if defined(__CUDA_ARCH__)
    return rsqrt(d);
#else
    return 1.0 / sqrt(d);
#endif
}

// Solve (Hs + diag(D2)) y = gs by Cholesky (D2 = the SQUARED damping diagonal); false when the matrix is not numerically
// positive definite.  The factor is kept with the INVERSE of its diagonal (L[c][c] holds 1 / l_cc), so the whole solve
// costs seven reciprocal square roots and no division: on the device this is a dependent chain on one thread.
PPCR_HD bool solve_damped(const double* Hs, const double* gs, const double* D2, double* y)
{
    double L[kNP][kNP];
    for (int r = 0; r < kNP; ++r)
        for (int c = 0; c < kNP; ++c) L[r][c] = Hs[r * kNP + c] + (r == c ? D2[r] : 0.0);
    for (int c = 0; c < kNP; ++c) {
        double d = L[c][c];
        for (int k = 0; k < c; ++k) d -= L[c][k] * L[c][k];
        if (!(d > 0.0)) return false;
        const double inv = inv_sqrt(d);
        L[c][c] = inv;
        for (int r = c + 1; r < kNP; ++r) {
            double s = L[r][c];
            for (int k = 0; k < c; ++k) s -= L[r][k] * L[c][k];
            L[r][c] = s * inv;
        }
    }
    double z[kNP];
    for (int r = 0; r < kNP; ++r) {
        double s = gs[r];
        for (int k = 0; k < r; ++k) s -= L[r][k] * z[k];
        z[r] = s * L[r][r];
    }
    for (int r = kNP - 1; r >= 0; --r) {
        double s = z[r];
        for (int k = r + 1; k < kNP; ++k) s -= L[k][r] * y[k];
        y[r] = s * L[r][r];
    }
    return true;
}

// Loads a fresh evaluation into the state: column-scaled Hs/gs, x_cost, gradient max-norm.
PPCR_HD void load_evaluation(PairState* s, const Expanded& ev, bool first)
{
    const double* H = ev.H;
    const double* g = ev.g;
    s->x_cost = ev.cost;
    double gm = 0.0;
    for (int p = 0; p < kNP; ++p) gm = fmax(gm, fabs(g[p]));
    s->grad_max = gm;  // max-norm of the UNSCALED gradient
    if (first) {       // Jacobi scaling is estimated once, at iteration zero: 1 / (1 + sqrt(|J_col|^2))
        for (int p = 0; p < kNP; ++p) s->scale[p] = 1.0 / (1.0 + sqrt(H[p * kNP + p]));
    }
    for (int r = 0; r < kNP; ++r) {
        s->gs[r] = g[r] * s->scale[r];
        for (int c = 0; c < kNP; ++c) s->Hs[r * kNP + c] = H[r * kNP + c] * s->scale[r] * s->scale[c];
    }
}

PPCR_HD double vec_norm7(const double* v)
{
    double n = 0.0;
    for (int p = 0; p < kNP; ++p) n += v[p] * v[p];
    return sqrt(n);
}

// FinalizeIteration + ComputeTrustRegionStep of the restated Ceres minimiser, cut at the linear solve so that the
// controller block can run the 7x7 Cholesky on several threads:
//   step_prepare  -> false: the minimiser stopped (s->termination says why); true: D2 holds the squared damping diagonal
//   [solve (Hs + diag(D)^2) y = gs]
//   step_complete -> 0: candidate ready in s->cand / s->pose_e (evaluate it next), 1: invalid step, prepare again,
//                    2: the minimiser stopped
PPCR_HD bool step_prepare(PairState* s, const Config* cfg, double* D)
{
    const double kMinRadius = 1e-32, kMinDiag = 1e-6, kMaxDiag = 1e32, kGradTol = 1e-10;
    if (s->step_ok) {
        ++s->successful;
        if (s->x_cost < s->minimum_cost) {
            s->minimum_cost = s->x_cost;
            for (int p = 0; p < kNP; ++p) s->best_x[p] = s->x[p];
        }
    }
    // The IterationCallback refreshes the weights at the state Ceres exposes to it: with update_state_every_iteration that is
    // the minimiser's `parameters_`, which only follows x when x_cost < minimum_cost (above) -- after an accepted
    // non-monotonic step it is still the lowest-cost iterate.  The next eval refreshes the weights there on the fly.
    pose_from_x(s->best_x, &s->pose_w);
    if (s->iteration >= cfg->max_lm_iterations) { s->termination = TERM_MAX_ITER; return false; }
    if (s->step_ok && s->grad_max <= kGradTol) { s->termination = TERM_GRADIENT_TOL; return false; }
    if (s->radius < kMinRadius) { s->termination = TERM_MIN_RADIUS; return false; }
    ++s->iteration;
    if (!s->reuse_diag) {
        for (int p = 0; p < kNP; ++p) s->diag[p] = fmin(fmax(s->Hs[p * kNP + p], kMinDiag), kMaxDiag);
    }
    const double inv_radius = 1.0 / s->radius;
    for (int p = 0; p < kNP; ++p) D[p] = s->diag[p] * inv_radius;  // (sqrt(diag / radius))^2, what the solve adds to the diagonal
    return true;
}

PPCR_HD int step_complete(PairState* s, const Config* cfg, bool valid, const double* y)
{
    (void)cfg;
    const int kMaxInvalid = 5;
    s->reuse_diag = 1;
    for (int p = 0; p < kNP && valid; ++p) valid = isfinite(y[p]);
    if (valid) {
        double lin = 0.0, quad = 0.0;
        for (int r = 0; r < kNP; ++r) s->step[r] = -y[r];
        for (int r = 0; r < kNP; ++r) {
            lin += s->step[r] * s->gs[r];
            double t = 0.0;
            for (int c = 0; c < kNP; ++c) t += s->Hs[r * kNP + c] * s->step[c];
            quad += s->step[r] * t;
        }
        s->model_change = -(lin + 0.5 * quad);
        valid = s->model_change > 0.0;
    }
    if (!valid) {
        if (++s->invalid >= kMaxInvalid) { s->termination = TERM_INVALID_STEPS; return 2; }
        s->radius /= s->decrease_factor;
        s->decrease_factor *= 2.0;
        s->step_ok = 0;
        s->min_iter_cost = fmin(s->min_iter_cost, s->x_cost);
        return 1;
    }
    s->invalid = 0;
    for (int p = 0; p < kNP; ++p) s->cand[p] = s->x[p] + s->step[p] * s->scale[p];
    pose_from_x(s->cand, &s->pose_e);
    return 0;
}

// the serial driver of the two halves (CPU build, and the reference the cooperative device driver must match)
PPCR_HD bool finalize_and_step(PairState* s, const Config* cfg)
{
    for (;;) {
        double D[kNP], y[kNP];
        if (!step_prepare(s, cfg, D)) return false;
        const bool valid = solve_damped(s->Hs, s->gs, D, y);
        const int r = step_complete(s, cfg, valid, y);
        if (r == 0) return true;
        if (r == 2) return false;
    }
}

PPCR_HD void lm_reset(PairState* s, const Config* cfg)
{
    for (int p = 0; p < kNP; ++p) {
        s->x[p] = cfg->x0[p];
        s->best_x[p] = cfg->x0[p];
        s->cand[p] = cfg->x0[p];
    }
    pose_from_x(s->x, &s->pose_e);  // iteration zero evaluates residuals and weights at the start pose
    pose_from_x(s->x, &s->pose_w);
    s->iteration = 0;
    s->invalid = 0;
    s->successful = 0;
    s->reuse_diag = 0;
    s->step_ok = 1;
    s->termination = -1;
    s->radius = 1e4;
    s->decrease_factor = 2.0;
    s->minimum_cost = 1.7976931348623157e308;
    s->initial_cost = 0.0;
    s->min_iter_cost = 0.0;
    s->x_cost = 0.0;
    s->K = 0;  // the search kernel accumulates the association size here
}

// After the first eval of an outer iteration (residuals and weights at x0).  true = LM continues with a
// trust-region step (step_prepare / solve / step_complete).
PPCR_HD bool lm_begin_eval(PairState* s, const Config* cfg, const Expanded& ev)
{
    (void)cfg;
    if (s->K == 0) {  // a Ceres problem without residual blocks: zero costs, pose untouched
        s->termination = TERM_NO_RESIDUALS;
        s->initial_cost = 0.0;
        s->min_iter_cost = 0.0;
        return false;
    }
    load_evaluation(s, ev, true);
    s->x_norm = vec_norm7(s->x);
    s->initial_cost = s->x_cost;
    s->min_iter_cost = s->x_cost;
    s->ev_min = s->ev_cur = s->ev_ref = s->ev_cand = s->x_cost;
    s->ev_acc_ref = 0.0;
    s->ev_acc_cand = 0.0;
    s->ev_nonmono = 0;
    s->step_ok = 1;
    return true;
}

// After an eval at the candidate (weights refreshed at the current iterate).  true = LM continues with a
// trust-region step.
PPCR_HD bool lm_continue_eval(PairState* s, const Config* cfg, const Expanded& ev)
{
    const double kParamTol = 1e-8, kMinRelDecrease = 1e-3, kMaxRadius = 1e16;
    const int kMaxNonmono = 5;
    const double cand_cost = ev.cost;
    double step_norm = 0.0;
    for (int p = 0; p < kNP; ++p) step_norm += (s->x[p] - s->cand[p]) * (s->x[p] - s->cand[p]);
    step_norm = sqrt(step_norm);
    if (step_norm <= kParamTol * (s->x_norm + kParamTol)) { s->termination = TERM_PARAMETER_TOL; return false; }
    if (fabs(s->x_cost - cand_cost) <= cfg->function_tolerance * s->x_cost) { s->termination = TERM_FUNCTION_TOL; return false; }

    const double rel = (s->ev_cur - cand_cost) / s->model_change;
    const double hist = (s->ev_ref - cand_cost) / (s->ev_acc_ref + s->model_change);
    const double quality = fmax(rel, hist);
    if (quality > kMinRelDecrease) {
        for (int p = 0; p < kNP; ++p) s->x[p] = s->cand[p];
        s->x_norm = vec_norm7(s->x);
        load_evaluation(s, ev, false);  // the same pass already produced the Jacobian at the candidate
        s->step_ok = 1;
        const double t = 2.0 * quality - 1.0;
        s->radius = fmin(kMaxRadius, s->radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
        s->decrease_factor = 2.0;
        s->reuse_diag = 0;
        // TrustRegionStepEvaluator::StepAccepted
        s->ev_cur = cand_cost;
        s->ev_acc_cand += s->model_change;
        s->ev_acc_ref += s->model_change;
        if (s->ev_cur < s->ev_min) {
            s->ev_min = s->ev_cur;
            s->ev_nonmono = 0;
            s->ev_cand = s->ev_cur;
            s->ev_acc_cand = 0.0;
        } else {
            ++s->ev_nonmono;
            if (s->ev_cur > s->ev_cand) {
                s->ev_cand = s->ev_cur;
                s->ev_acc_cand = 0.0;
            }
        }
        if (s->ev_nonmono == kMaxNonmono) {
            s->ev_ref = s->ev_cand;
            s->ev_acc_ref = s->ev_acc_cand;
        }
        s->min_iter_cost = fmin(s->min_iter_cost, s->x_cost);
    } else {
        s->step_ok = 0;
        s->radius /= s->decrease_factor;
        s->decrease_factor *= 2.0;
        s->reuse_diag = 1;
        s->min_iter_cost = fmin(s->min_iter_cost, cand_cost);
    }
    return true;
}

// hasConverged(), src/prob_point_cloud_registration.cc:138-158 (mutating).
PPCR_HD bool has_converged(PairState* s, const Config* cfg)
{
    if (s->current_iteration == cfg->n_iter) return true;
    if (s->cost_drop < cfg->cost_drop_thresh) {
        if (static_cast<double>(s->num_unuseful) > cfg->n_cost_drop_it) return true;
        ++s->num_unuseful;
    } else {
        s->num_unuseful = 0;
    }
    return false;
}

// The tail of one outer iteration, src/prob_point_cloud_registration.cc:101-130, once the inner LM stopped.
PPCR_HD void outer_finish(PairState* s, const Config* cfg, double* history, IterStats* stats, int32_t max_hist)
{
    for (int p = 0; p < kNP; ++p) s->x[p] = s->best_x[p];  // the minimiser hands back its lowest-cost iterate
    matrix_from_x(s->x, s->dT);
    if (s->current_iteration > 0) {
        matmul4(s->dT, s->T_total, s->T_total);
    } else {
        for (int k = 0; k < 16; ++k) s->T_total[k] = s->dT[k];
    }
    s->cost_drop = (s->initial_cost - s->min_iter_cost) / s->initial_cost;  // NaN for 0/0, like the reference
    if (s->current_iteration < max_hist) {
        for (int k = 0; k < 16; ++k) {
            history[16 * s->current_iteration + k] = s->T_total[k];         // accumulated pose, registration.cc:101-107
            history[16 * (max_hist + s->current_iteration) + k] = s->dT[k]; // the increment the clouds are moved by, :110-112
        }
        IterStats* st = stats + s->current_iteration;
        st->initial_cost = s->initial_cost;
        st->final_cost = s->min_iter_cost;
        st->cost_drop = s->cost_drop;
        st->n_correspondences = s->K;
        st->lm_iterations = s->iteration;
        st->num_successful_steps = s->successful;
    }
    s->K_total += s->K;
    s->apply_dT = 1;  // the next search (or the epilogue of align()) moves the cloud by dT first
    s->search_cursor = 0;
    ++s->current_iteration;
    s->phase = has_converged(s, cfg) ? PH_DONE : PH_SEARCH;
    if (s->phase == PH_SEARCH) lm_reset(s, cfg);
}

// Pose the evaluation that has just finished took its residuals at (the moments are expanded around it).
PPCR_HD const double* evaluated_at(const PairState* s) { return s->phase == PH_SEARCH ? s->x : s->cand; }

// First half of a controller invocation: digest the evaluation.  true = a trust-region step follows.
PPCR_HD bool controller_begin(PairState* s, const Config* cfg, const Expanded& ev)
{
    ++s->ticks;
    bool go;
    if (s->phase == PH_SEARCH) {
        s->apply_dT = 0;  // the search that fed this evaluation has moved the cloud
        go = lm_begin_eval(s, cfg, ev);
        s->phase = PH_LM;
    } else {
        go = lm_continue_eval(s, cfg, ev);
    }
    return go;
}

// One controller invocation after an eval, serial form (CPU build; the device runs the same pieces with the moment
// expansion and the linear solve spread over threads -- k_evalctl).
PPCR_HD void controller_tick(PairState* s, const Config* cfg, const double* S, double* history, IterStats* stats,
                             int32_t max_hist)
{
    if (s->phase == PH_DONE) return;
    Expanded ev;
    expand_moments(S, evaluated_at(s), &ev);
    bool go = controller_begin(s, cfg, ev);
    if (go) go = finalize_and_step(s, cfg);
    if (!go) outer_finish(s, cfg, history, stats, max_hist);
}

PPCR_HD void state_init(PairState* s, const Config* cfg)
{
    s->cost_drop = 0.0;
    s->current_iteration = 0;
    s->num_unuseful = 0;
    s->apply_dT = 0;
    s->K = 0;
    s->K_total = 0;
    s->exchange_cycles = 0;
    s->ticks = 0;
    s->evals = 0;
    s->error = 0;
    s->row_overflow = 0;
    s->search_cursor = 0;
    s->eval_ticket = 0;
    for (int k = 0; k < 16; ++k) {
        s->T_total[k] = (k % 5 == 0) ? 1.0 : 0.0;
        s->dT[k] = (k % 5 == 0) ? 1.0 : 0.0;
    }
    lm_reset(s, cfg);
    s->phase = PH_DONE;  // nothing runs until align_begin() has made the first hasConverged() test
}

// The first `while (!hasConverged())` test of align() (:65).  Later tests happen in outer_finish().
PPCR_HD void align_begin(PairState* s, const Config* cfg)
{
    s->apply_dT = 0;
    s->search_cursor = 0;
    if (has_converged(s, cfg)) {
        s->phase = PH_DONE;
    } else {
        lm_reset(s, cfg);
        s->phase = PH_SEARCH;
    }
}

}  // namespace ppcr
#endif
