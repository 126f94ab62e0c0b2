// ppcr_eval.h -- per-source-point arithmetic of the fused "weights + normal equations" pass.
//
// One source point i with its k_i associated target points y_j produces
//   w_ij  : probabilistic_weights.hpp:48-105, evaluated from the residuals at pose_w
//           (the pose of the last WeightUpdaterCallback, weight_updater_callback.hpp:36-63), and
//   r_ij  : error_term.hpp:21-37 at pose_e (the pose Ceres is evaluating),
// folded into the 24 moments of ppcr_lm.h.  The row softmax of the reference (max-subtracted log-sum-exp over the
// row, hpp:77-99) is computed in ONE pass with a running maximum, so neither the weights nor the residuals are
// ever stored: the pass reads 12 bytes per correspondence and 20 bytes per source point.
//
// Plain C++ qualified PPCR_HD: the eval kernel and the CPU tests of the host logic share this source.
#ifndef PPCR_EVAL_H
#define PPCR_EVAL_H

#include "ppcr_lm.h"

namespace ppcr {

struct WeightCfg {
    double dof;        // v
    double t_exponent; // -(v + 3) / 2        (probabilistic_weights.hpp:37, DIMENSIONS = 3)
    double dof_plus_d; // v + 3               (numerator of the expected weight, hpp:73)
    double inv_dof;    // 1 / v is NOT used for the log argument (the reference divides); kept for the fp32 path
    int32_t is_normal; // v == +inf: Gaussian model, w = softmax(-r^2/2)
    int32_t pad;
};

PPCR_HD WeightCfg make_weight_cfg(double dof, int dimension = 3)
{
    const double dim = static_cast<double>(dimension);  // DIMENSIONS = 3 in production (iteration.hpp:17,29)
    WeightCfg w;
    w.is_normal = !(dof < 1.7976931348623157e308);
    w.dof = dof;
    w.t_exponent = w.is_normal ? 0.0 : -(dof + dim) / 2.0;
    w.dof_plus_d = dof + dim;
    w.inv_dof = w.is_normal ? 0.0 : 1.0 / dof;
    w.pad = 0;
    return w;
}

// log-probability up to the row-constant normaliser (which cancels in the softmax) and expected weight
template <bool kFast>
PPCR_HD void log_prob(const WeightCfg& wc, double r2, double* lp, double* expected)
{
    if (wc.is_normal) {
        *lp = -r2 / 2.0;
        *expected = 1.0;
    } else if (kFast) {
        const float z = static_cast<float>(r2 / wc.dof);
        *lp = static_cast<double>(static_cast<float>(wc.t_exponent) * log1pf(z));
        *expected = static_cast<double>(static_cast<float>(wc.dof_plus_d) / (static_cast<float>(wc.dof) + static_cast<float>(r2)));
    } else {
        *lp = wc.t_exponent * log1p(r2 / wc.dof);
        *expected = wc.dof_plus_d / (wc.dof + r2);
    }
}

template <bool kFast>
PPCR_HD double exp_diff(double d)
{
    if (kFast) return static_cast<double>(expf(static_cast<float>(d)));
    return exp(d);
}

struct RowAcc {  // running softmax state of one source row
    double m;    // running max of the log-probabilities
    double a0;   // sum exp(l - m)
    double a1;   // sum exp(l - m) e
    double ar[3];// sum exp(l - m) e r
    double ac;   // sum exp(l - m) e |r|^2
};

PPCR_HD void row_begin(RowAcc* a)
{
    a->m = -1.7976931348623157e308;
    a->a0 = a->a1 = a->ac = 0.0;
    a->ar[0] = a->ar[1] = a->ar[2] = 0.0;
}

// one correspondence: target point (yx,yy,yz); pe = R_e x + t_e, pw = R_w x + t_w already computed for the row
template <bool kFast>
PPCR_HD void row_add(RowAcc* a, const WeightCfg& wc, double yx, double yy, double yz, const double* pe, const double* pw)
{
    const double wx = yx - pw[0], wy = yy - pw[1], wz = yz - pw[2];
    const double r2w = wx * wx + wy * wy + wz * wz;  // squared error the callback hands to updateWeights
    double lp, ex;
    log_prob<kFast>(wc, r2w, &lp, &ex);
    const double ex_ = ex;
    const double rx = yx - pe[0], ry = yy - pe[1], rz = yz - pe[2];
    const double r2e = rx * rx + ry * ry + rz * rz;
    if (lp > a->m) {  // new row maximum: rescale what has been accumulated so far
        const double sc = exp_diff<kFast>(a->m - lp);
        a->a0 *= sc;
        a->a1 *= sc;
        a->ar[0] *= sc;
        a->ar[1] *= sc;
        a->ar[2] *= sc;
        a->ac *= sc;
        a->m = lp;
    }
    const double p = exp_diff<kFast>(lp - a->m);
    const double pe_w = p * ex_;
    a->a0 += p;
    a->a1 += pe_w;
    a->ar[0] += pe_w * rx;
    a->ar[1] += pe_w * ry;
    a->ar[2] += pe_w * rz;
    a->ac += pe_w * r2e;
}

// fold a finished row into the 24 moments; (sx,sy,sz) is the source point in double
PPCR_HD void row_end(const RowAcc* a, double sx, double sy, double sz, double* acc)
{
    const double inv = 1.0 / a->a0;
    const double W = a->a1 * inv;
    const double rho[3] = {a->ar[0] * inv, a->ar[1] * inv, a->ar[2] * inv};
    acc[M_S0] += W;
    acc[M_S1 + 0] += W * sx;
    acc[M_S1 + 1] += W * sy;
    acc[M_S1 + 2] += W * sz;
    acc[M_S2 + 0] += W * sx * sx;
    acc[M_S2 + 1] += W * sx * sy;
    acc[M_S2 + 2] += W * sx * sz;
    acc[M_S2 + 3] += W * sy * sy;
    acc[M_S2 + 4] += W * sy * sz;
    acc[M_S2 + 5] += W * sz * sz;
    acc[M_SR + 0] += rho[0];
    acc[M_SR + 1] += rho[1];
    acc[M_SR + 2] += rho[2];
    acc[M_C + 0] += sx * rho[0];
    acc[M_C + 1] += sx * rho[1];
    acc[M_C + 2] += sx * rho[2];
    acc[M_C + 3] += sy * rho[0];
    acc[M_C + 4] += sy * rho[1];
    acc[M_C + 5] += sy * rho[2];
    acc[M_C + 6] += sz * rho[0];
    acc[M_C + 7] += sz * rho[1];
    acc[M_C + 8] += sz * rho[2];
    acc[M_COST] += 0.5 * a->ac * inv;
    acc[M_ROWS] += 1.0;
}

PPCR_HD void apply_pose(const Pose& p, double sx, double sy, double sz, double* out)
{
    out[0] = p.R[0] * sx + p.R[1] * sy + p.R[2] * sz + p.t[0];
    out[1] = p.R[3] * sx + p.R[4] * sy + p.R[5] * sz + p.t[1];
    out[2] = p.R[6] * sx + p.R[7] * sy + p.R[8] * sz + p.t[2];
}

// weight of one correspondence once the row statistics (m, a0) are known -- used only when the weights
// themselves are requested (parity dumps); the solver never materialises them.
template <bool kFast>
PPCR_HD double finished_weight(const RowAcc* a, const WeightCfg& wc, double yx, double yy, double yz, const double* pw)
{
    const double wx = yx - pw[0], wy = yy - pw[1], wz = yz - pw[2];
    const double r2w = wx * wx + wy * wy + wz * wz;
    double lp, ex;
    log_prob<kFast>(wc, r2w, &lp, &ex);
    return exp_diff<kFast>(lp - a->m) / a->a0 * ex;
}

}  // namespace ppcr
#endif
